"""-m gpu: parity at BASELINE.json's (scaled) configuration sizes, against the REFERENCE code.

The indexes are the ones bench.py uses, cached under .cache/ by `bench.py` / the dev container
(this repo's builder for the GPU, the reference's own builder+serializer for oracle/_ref). A test
skips when its cache files did not travel. Texts and patterns are regenerated deterministically.

  * a parity sample (first 1500 + 1500 random patterns): every range, offset and occurrence position
    equals what the reference's locate_all returns, in its order;
  * the whole batch: n_occ == hi-lo+1 per pattern, offsets are its prefix sums, count() and locate()
    agree, and an on-device digest of all occurrences equals the digest of the D2H copy.
"""
import os

import numpy as np
import pytest

from conftest import rib, ob, ROOT

pytestmark = pytest.mark.gpu
CACHE = os.path.join(ROOT, ".cache")


def _workload(name):
    import bench
    return bench.WORKLOADS[name]


@pytest.mark.parametrize("name", ["c2", "c5s", "c3s", "c4s"])
def test_scaled_config_against_reference(name):
    rib_path = os.path.join(CACHE, name + ".rib")
    ref_path = os.path.join(CACHE, name + ".ref.ri")
    if not (os.path.exists(rib_path) and os.path.exists(ref_path) and ob.have_ref()):
        pytest.skip("cached indexes for %s not present" % name)
    kind, n, p0, p1, tseed, N, m, pseed, limit, desc = _workload(name)
    text = rib.gen_text(kind, n, p0, p1, tseed)
    patt = rib.gen_patterns(text, N, m, pseed, limit)
    del text
    host = rib.HostIndex.load(rib_path)
    ref = ob.RefIndex.load(ref_path)
    assert (host.n, host.r) == (ref.n, ref.r)
    gpu = rib.GpuIndex(host)
    # ---- parity sample vs the reference's own code ----
    rng = np.random.default_rng(7)
    pick = np.unique(np.concatenate([np.arange(min(N, 1500)), rng.integers(0, N, size=1500)]))
    sub = patt.reshape(N, m)[pick].reshape(-1).copy()
    S = pick.size
    threads = os.cpu_count() or 1
    if name == "c4s":   # ri-count config: ranges only (locating ~10^2 occ/read x 1e6 reads is not its job)
        elo, ehi, _ = ref.count(sub, S, m, threads=threads)
        lo, hi = gpu.count(sub, S, m)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi)
    else:
        elo, ehi, eoff, eocc, _ = ref.locate(sub, S, m, threads=threads)
        lo, hi, off, occ = gpu.locate(sub, S, m)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi)
        assert np.array_equal(off, eoff)
        assert np.array_equal(occ, eocc)
        assert eocc.size > 0
    # ---- whole batch: size-independent properties ----
    lo, hi = gpu.count(patt, N, m)
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
    assert np.array_equal(lo[pick], elo) and np.array_equal(hi[pick], ehi)
    if name != "c4s":
        torch = pytest.importorskip("torch")
        lo2, hi2, off, occ = gpu.locate(patt, N, m)
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
        assert np.array_equal(np.diff(off), nocc) and occ.size == int(nocc.sum())
        assert int(occ.max()) <= n - m
        from rindex_b200._gpu import digest_host
        d_occ = torch.from_numpy(occ.view(np.int64)).to("cuda:0")
        assert gpu.digest_dev(d_occ.data_ptr(), d_occ.numel()) == digest_host(occ)
        # positions of a pattern are distinct
        for p in rng.integers(0, N, size=50):
            o = occ[int(off[p]):int(off[p + 1])]
            assert np.unique(o).size == o.size


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_full_size_config_self_check(name):
    """BASELINE.json configs 3 and 5 at FULL size (1 GB sigma=96 text; 4 GB DNA, n = 4.0e9: 7% under 2^32, still 32-bit words).
    No CPU oracle finishes at this size, so the device runs the reference's own -c self-check
    (ri-locate.cpp:156-190) on the located output: brute-force occurrence counts from the text (hash join)
    equal hi-lo+1 for every pattern, text[o, o+m) equals the pattern for every located o, and the positions of
    a pattern are distinct. The index comes from .cache/ (built in the dev container: 5 / 29 minutes)."""
    rib_path = os.path.join(CACHE, name + ".rib")
    if not os.path.exists(rib_path):
        pytest.skip("cached index for %s not present" % name)
    kind, n, p0, p1, tseed, N, m, pseed, limit, desc = _workload(name)
    N = 20_000
    text = rib.gen_text(kind, n, p0, p1, tseed)
    patt = rib.gen_patterns(text, N, m, pseed, limit)
    host = rib.HostIndex.load(rib_path)
    assert host.n == n + 1
    gpu = rib.GpuIndex(host)
    assert gpu.info.words32 == (1 if n + 1 < 2**32 - 1 else 0) and gpu.info.seed_jump == 64
    gpu.text_attach(text)
    del text
    lo, hi, off, occ, rep = gpu.locate_ex(patt, N, m, rib.LOCATE_SORT | rib.LOCATE_CHECK)
    assert rep.patterns_checked == N and rep.clean, rep.as_dict()
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
    assert (nocc > 0).all() and np.array_equal(np.diff(off), nocc) and occ.size == int(nocc.sum())
    lo2, hi2 = gpu.count(patt, N, m)
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    # the unsorted locate_all order holds the same multiset per pattern (spot check)
    lo3, hi3, off3, occ3 = gpu.locate(patt[: 50 * m], 50, m)
    for p in range(50):
        assert np.array_equal(np.sort(occ3[int(off3[p]):int(off3[p + 1])]), occ[int(off[p]):int(off[p + 1])])
