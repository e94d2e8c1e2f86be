"""-m gpu: CUDA path vs the committed golden fixtures (generated from the reference's own code,
tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import rib, GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [4, 8, 16])
@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")))
def test_gpu_vs_golden(fname, K):
    g = np.load(os.path.join(GOLDEN, fname))
    host = rib.HostIndex.from_text(g["text"])
    gpu = rib.GpuIndex(host, runs_per_block=K)
    N, m = int(g["N"]), int(g["m"])
    lo, hi = gpu.count(g["patterns"], N, m)
    assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"])
    lo, hi, off, occ = gpu.locate(g["patterns"], N, m)
    assert np.array_equal(lo, g["lo"]) and np.array_equal(hi, g["hi"])
    assert np.array_equal(off, g["occ_offsets"])
    assert hashlib.sha256(occ.tobytes()).hexdigest() == str(g["occ_sha256"])
