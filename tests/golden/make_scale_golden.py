#!/usr/bin/env python
"""Reference-generated ANSWERS for BASELINE.json's configs at their scaled and FULL sizes, small enough to commit.

Run in the dev container (needs /root/reference for oracle/_ref; C5 at full size: ~30 min, ~45 GB RAM):

    python tests/golden/make_scale_golden.py [c2 c5s c3s c4s c3 c5]

For every workload of bench.py named on the command line (default: all six):
  * the text and the patterns are regenerated from bench.WORKLOADS' seeds (rank 0's pattern file);
  * the index is built by the REFERENCE's own constructor (r_index.hpp:42-150, suffix array through oracle/sdsl_shim)
    — or loaded from .cache/<name>.ref.ri when an earlier run of this script or of bench.py left it there;
  * a parity sample of patterns (the first K and K random ones; K = 1500 for the scaled configs, 150 for the
    full-size ones whose patterns have ~10^4 occurrences each) is run through the reference's own count() /
    locate_all() (oracle/ref_driver.cpp);
  * tests/golden/scale/<name>.npz keeps: the sample's pattern indices, lo, hi, per-pattern (sum, index-weighted
    sum) digests of the occurrences IN locate_all ORDER (r_index.hpp:340-351) and the sha256 of the whole sample's
    occurrence array; n, r and the sha256 of each logical array of the reference-built index (F, run heads, run
    lengths, samples_last, pred positions, pred_to_run — ref_extract), so the GPU box can check that the index it
    builds with the prefix-free-parsing builder IS the reference's index without either cache file.

tests/test_gpu_scale.py consumes these files; nothing else does.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

rib = ge.load_package()
ob = ge.load_oracle()

ARRAYS = ("F", "run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run")
COUNT_ONLY = {"c4s"}          # ri-count configs: ranges only
SAMPLE_K = {"c3": 150, "c5": 150}


def sample_indices(N, K):
    rng = np.random.default_rng(7)
    return np.unique(np.concatenate([np.arange(min(N, K)), rng.integers(0, N, size=K)])).astype(np.int64)


def per_pattern_digests(off, occ):
    """(sum, sum of v * (k+1)) mod 2^64 over each pattern's occurrences, k = position inside the pattern's list."""
    S = off.size - 1
    d = np.zeros((S, 2), dtype=np.uint64)
    for p in range(S):
        o = occ[int(off[p]):int(off[p + 1])]
        w = np.arange(1, o.size + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            d[p, 0] = np.add.reduce(o, dtype=np.uint64)
            d[p, 1] = np.add.reduce(o * w, dtype=np.uint64)
    return d


def make(name):
    kind, n, p0, p1, tseed, N, m, pseed, limit, desc = bench.WORKLOADS[name]
    t0 = time.time()
    text = rib.gen_text(kind, n, p0, p1, tseed)
    patt = rib.gen_patterns(text, N, m, pseed, limit)
    rpath = os.path.join(ROOT, ".cache", name + ".ref.ri")
    os.makedirs(os.path.dirname(rpath), exist_ok=True)
    if os.path.exists(rpath):
        ref = ob.RefIndex.load(rpath)
        built = "loaded .cache/%s.ref.ri (written by the reference's serialize())" % name
    else:
        print("[%s] building the REFERENCE index (n = %d) ..." % (name, n), flush=True)
        ref = ob.RefIndex.from_text(text)
        ref.save(rpath)
        built = "reference constructor, %.0f s" % (time.time() - t0)
    del text
    K = SAMPLE_K.get(name, 1500)
    pick = sample_indices(N, K)
    sub = patt.reshape(N, m)[pick].reshape(-1).copy()
    S = pick.size
    threads = os.cpu_count() or 1
    out = {"name": name, "n": np.uint64(ref.n), "r": np.uint64(ref.r), "N": np.int64(N), "m": np.int64(m), "pick": pick,
           "patt_sha256": hashlib.sha256(patt.tobytes()).hexdigest(), "built": built}
    if name in COUNT_ONLY:
        lo, hi, _ = ref.count(sub, S, m, threads=threads)
        out.update(lo=lo, hi=hi)
    else:
        lo, hi, off, occ, _ = ref.locate(sub, S, m, threads=threads)
        out.update(lo=lo, hi=hi, occ_total=np.uint64(occ.size), occ_sha256=hashlib.sha256(occ.tobytes()).hexdigest(),
                   occ_digests=per_pattern_digests(off, occ))
    ex = ref.extract()
    for k in ARRAYS:
        out["sha256_" + k] = hashlib.sha256(np.ascontiguousarray(ex[k]).tobytes()).hexdigest()
    os.makedirs(os.path.join(HERE, "scale"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "scale", "%s.npz" % name), **out)
    print("[%s] n=%d r=%d sample=%d patterns%s  (%.0f s)" % (
        name, ref.n, ref.r, S, "" if name in COUNT_ONLY else ", %d occurrences" % int(out["occ_total"]), time.time() - t0),
        flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["c2", "c5s", "c3s", "c4s", "c3", "c5"]):
        make(nm)
