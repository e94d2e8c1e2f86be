"""-m gpu: BASELINE.json configs[0] — the reference's bundled world_leaders text and pattern file.
Expected values were produced by the REFERENCE code (tests/golden/make_config1.py) and agree with
SURVEY.md §4's brute-force known answer occ_t = 29,781,174. The index (built by this repo's builder,
verified equal to the reference-built one) and the pattern file are cached under .cache/ by that
script; the GPU box has no /root/reference, so the test skips if the cache did not travel."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import rib, ROOT, GOLDEN

pytestmark = pytest.mark.gpu
IDX = os.path.join(ROOT, ".cache", "world_leaders.rib")
PATT = os.path.join(ROOT, ".cache", "world_leaders_1000_8.patt")


@pytest.mark.skipif(not (os.path.exists(IDX) and os.path.exists(PATT)), reason=".cache/world_leaders.* not present")
def test_config1_world_leaders():
    g = json.load(open(os.path.join(GOLDEN, "config1_world_leaders.json")))
    N, m, patt = rib.parse_pattern_file(open(PATT, "rb").read())
    assert (N, m) == (g["N"], g["m"]) == (1000, 8)
    host = rib.HostIndex.load(IDX)
    assert (host.n, host.r) == (g["n"], g["r"])
    gpu = rib.GpuIndex(host)
    lo, hi = gpu.count(patt, N, m)
    nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
    assert int(nocc.sum()) == g["occ_t"] == 29_781_174          # what ri-count prints as occ_t
    assert hashlib.sha256(nocc.astype("<u8").tobytes()).hexdigest() == g["counts_sha256"]
    assert hashlib.sha256(lo.tobytes()).hexdigest() == g["lo_sha256"]
    assert hashlib.sha256(hi.tobytes()).hexdigest() == g["hi_sha256"]
    lo2, hi2, off, occ = gpu.locate(patt, N, m)
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and occ.size == g["occ_t"]
    assert hashlib.sha256(occ.tobytes()).hexdigest() == g["occ_sha256"]   # every position, in locate_all order
    t = gpu.timing()
    print("config1: r=%d chains=%d search=%.3fms expand=%.3fms" % (host.r, t["chains"], t["search_ms"], t["expand_ms"]))
