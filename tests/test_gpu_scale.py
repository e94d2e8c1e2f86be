"""-m gpu: parity at BASELINE.json's configuration sizes — scaled (C2, C5s, C3s, C4s) and FULL (C3: 1 GB sigma=96;
C5: 4 GB DNA, n = 4.0e9) — against ANSWERS PRODUCED BY THE REFERENCE's own code.

tests/golden/scale/<name>.npz (written by tests/golden/make_scale_golden.py in the dev container, where the
reference's constructor and locate_all ran through oracle/_ref) holds, per workload: the sample's pattern indices,
the reference's lo / hi for them, per-pattern digests and the SHA-256 of the sample's occurrences in locate_all
order (r_index.hpp:340-351), and the SHA-256 of each logical array of the reference-BUILT index. Nothing here
needs /root/reference or a cached index: texts and patterns are regenerated from their seeds, and an index that is
not under .cache/ is rebuilt on the spot with the prefix-free-parsing builder (1 GB ~ 6 s, 4 GB ~ 25 s). No skips.

  * the index in use (cached or rebuilt) IS the reference-built index: array by array, by hash;
  * the parity sample: every range, and every occurrence position in the reference's order, equal the reference's;
  * the whole batch (or its first 20 k patterns at full size): n_occ == hi-lo+1, offsets are prefix sums, count()
    and locate() agree, positions of a pattern are distinct; at full size the device also runs ri-locate's own -c
    self-check over every located position (ri-locate.cpp:156-190).

C4 at full size (10 GB) has no reference answers: the reference's constructor needs a 10^10-entry 64-bit suffix
array (80 GB) and this container has 62 GB; its witness stays the device self-check (tools/c4_selfcheck.py).
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import rib, ROOT, GOLDEN

pytestmark = pytest.mark.gpu
CACHE = os.path.join(ROOT, ".cache")
ARRAYS = ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run")
COUNT_ONLY = {"c4s"}
FULL_SIZE = {"c3", "c5"}


def _workload(name):
    import bench
    return bench.WORKLOADS[name]


def _digests(off, occ):
    S = off.size - 1
    d = np.zeros((S, 2), dtype=np.uint64)
    for p in range(S):
        o = occ[int(off[p]):int(off[p + 1])]
        w = np.arange(1, o.size + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            d[p, 0] = np.add.reduce(o, dtype=np.uint64)
            d[p, 1] = np.add.reduce(o * w, dtype=np.uint64)
    return d


@pytest.mark.parametrize("name", ["c2", "c5s", "c3s", "c4s", "c3", "c5"])
def test_config_against_reference_answers(name):
    g = np.load(os.path.join(GOLDEN, "scale", "%s.npz" % name))
    kind, n, p0, p1, tseed, N, m, pseed, limit, desc = _workload(name)
    assert (int(g["N"]), int(g["m"]), int(g["n"])) == (N, m, n + 1)
    text = rib.gen_text(kind, n, p0, p1, tseed)
    patt = rib.gen_patterns(text, N, m, pseed, limit)
    assert hashlib.sha256(patt.tobytes()).hexdigest() == str(g["patt_sha256"])   # the very patterns the reference answered
    rib_path = os.path.join(CACHE, name + ".rib")
    host = rib.HostIndex.load(rib_path) if os.path.exists(rib_path) else rib.HostIndex.from_text_auto(text)
    assert (host.n, host.r) == (int(g["n"]), int(g["r"]))
    a = host.arrays()
    for k in ARRAYS:   # the index in use is the index the reference's constructor built
        assert hashlib.sha256(np.ascontiguousarray(a[k]).tobytes()).hexdigest() == str(g["sha256_" + k]), k
    gpu = rib.GpuIndex(host)
    # ---- parity sample vs the reference's answers ----
    pick = g["pick"]
    sub = patt.reshape(N, m)[pick].reshape(-1).copy()
    S = pick.size
    if name in COUNT_ONLY:   # ri-count config: ranges only
        lo, hi = gpu.count(sub, S, m)
        assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"])
    else:
        lo, hi, off, occ = gpu.locate(sub, S, m)
        assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"])
        assert occ.size == int(g["occ_total"]) > 0
        d = _digests(off, occ)
        bad = np.nonzero((d != g["occ_digests"]).any(axis=1))[0]
        assert bad.size == 0, "patterns %s differ from the reference's locate_all output" % pick[bad[:5]]
        assert hashlib.sha256(occ.tobytes()).hexdigest() == str(g["occ_sha256"])   # every position, in the reference's order
    # ---- the batch: size-independent properties ----
    NB = min(N, 20_000) if name in FULL_SIZE else N
    batch = patt[: NB * m]
    lo, hi = gpu.count(batch, NB, m)
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
    inb = pick[pick < NB]
    assert np.array_equal(lo[inb], g["lo"][: inb.size]) and np.array_equal(hi[inb], g["hi"][: inb.size])
    if name in COUNT_ONLY:
        return
    if name in FULL_SIZE:
        # ri-locate -c on the device over all of the batch's occurrences: brute-force counts from the text (hash join)
        # equal hi-lo+1 for every pattern, text[o, o+m) equals the pattern for every located o, positions distinct
        assert gpu.info.words32 == (1 if n + 1 < 2**32 - 1 else 0) and gpu.info.seed_jump == 128
        gpu.text_attach(text)
        del text
        lo2, hi2, off, occ, rep = gpu.locate_ex(batch, NB, m, rib.LOCATE_SORT | rib.LOCATE_CHECK)
        assert rep.patterns_checked == NB and rep.clean, rep.as_dict()
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
        assert (nocc > 0).all() and np.array_equal(np.diff(off), nocc) and occ.size == int(nocc.sum())
        return
    del text
    torch = pytest.importorskip("torch")
    lo2, hi2, off, occ = gpu.locate(batch, NB, m)
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    assert np.array_equal(np.diff(off), nocc) and occ.size == int(nocc.sum())
    assert int(occ.max()) <= n - m
    from rindex_b200._gpu import digest_host
    d_occ = torch.from_numpy(occ.view(np.int64)).to("cuda:0")
    assert gpu.digest_dev(d_occ.data_ptr(), d_occ.numel()) == digest_host(occ)
    rng = np.random.default_rng(7)
    for p in rng.integers(0, NB, size=50):   # positions of a pattern are distinct
        o = occ[int(off[p]):int(off[p + 1])]
        assert np.unique(o).size == o.size
