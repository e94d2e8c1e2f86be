"""Test plumbing: markers, package loading, checker builds, shared fixtures.

Tiers:
  -m "not gpu"  CPU only: oracle vs golden vectors / brute force, host builder, flatten step via the
                scalar test double (tests/support/flat_check.cpp), C-ABI symbol export, CLI text.
  -m gpu        parity tests proper: CUDA path (through the C ABI) vs the oracle.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402

rib = ge.load_package()
ob = ge.load_oracle()

SUPPORT = os.path.join(ROOT, "tests", "support")
FLATCHECK_SO = os.path.join(SUPPORT, "libflat_check.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: multi-second CPU test")


def _build_flatcheck():
    src = os.path.join(SUPPORT, "flat_check.cpp")
    dep = os.path.join(ROOT, "r-index_b200", "csrc", "flat_layout.hpp")
    if (not os.path.exists(FLATCHECK_SO) or os.path.getmtime(FLATCHECK_SO) < max(os.path.getmtime(src), os.path.getmtime(dep))):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-o", FLATCHECK_SO, src], check=True)


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    rib.build_host()
    rib.build_gpu()   # nvcc cross-compiles without a GPU; no-op when librindex_gpu.so is up to date
    ob.build()
    _build_flatcheck()


class FlatCheck:
    """ctypes wrapper of the scalar test double over the flattened arrays."""

    def __init__(self, host_index, K=16, lf_log2=0, phi_log2=0, jump=0, force_wide=False, seed_jump=0):
        _build_flatcheck()
        self.lib = ctypes.CDLL(FLATCHECK_SO)
        self.lib.fc_create.restype = ctypes.c_void_p
        self.lib.fc_create.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int), ctypes.c_uint32]
        self.lib.fc_jump.restype = ctypes.c_uint64
        self.lib.fc_jump.argtypes = [ctypes.c_void_p]
        self.lib.fc_pieces.restype = ctypes.c_uint64
        self.lib.fc_pieces.argtypes = [ctypes.c_void_p]
        self.lib.fc_w32.argtypes = [ctypes.c_void_p]
        for fn in ("fc_seed_jump", "fc_seed_pieces"):
            getattr(self.lib, fn).restype = ctypes.c_uint64
            getattr(self.lib, fn).argtypes = [ctypes.c_void_p]
        self.lib.fc_check_seed.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        self.lib.fc_check_jump.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        self.lib.fc_destroy.argtypes = [ctypes.c_void_p]
        self.lib.fc_count.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_uint64] * 2 + [ctypes.c_void_p] * 2
        self.lib.fc_locate.restype = ctypes.c_uint64
        self.lib.fc_locate.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_uint64] * 2 + [ctypes.c_void_p] * 4
        rc = ctypes.c_int(0)
        if isinstance(host_index, dict):
            from rindex_b200._gpu import view_from_arrays
            view, self._keep = view_from_arrays(host_index)
        else:
            view, self._keep = host_index.view, host_index
        self.h = self.lib.fc_create(ctypes.byref(view), K, lf_log2, phi_log2, jump, int(force_wide), ctypes.byref(rc), seed_jump)
        self.rc = rc.value
        self.w32 = bool(self.lib.fc_w32(self.h)) if self.h else None
        self.jump = int(self.lib.fc_jump(self.h)) if self.h else 0
        self.phi_packed = bool(self.lib.fc_phi_packed(self.h)) if self.h else False
        self.seed_jump = int(self.lib.fc_seed_jump(self.h)) if self.h else 0

    def count(self, patt, N, m):
        p = np.ascontiguousarray(patt, dtype=np.uint8)
        lo = np.zeros(N, dtype=np.uint64)
        hi = np.zeros(N, dtype=np.uint64)
        self.lib.fc_count(self.h, p.ctypes.data, N, m, lo.ctypes.data, hi.ctypes.data)
        return lo, hi

    def locate(self, patt, N, m):
        p = np.ascontiguousarray(patt, dtype=np.uint8)
        lo, hi = self.count(p, N, m)
        nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
        off = np.zeros(N + 1, dtype=np.uint64)
        np.cumsum(nocc, out=off[1:])
        occ = np.full(int(off[-1]), np.uint64(2**64 - 1), dtype=np.uint64)
        chains = self.lib.fc_locate(self.h, p.ctypes.data, N, m, lo.ctypes.data, hi.ctypes.data, off.ctypes.data,
                                    occ.ctypes.data)
        return lo, hi, off, occ, int(chains)

    def __del__(self):
        try:
            if self.h:
                self.lib.fc_destroy(self.h)
        except Exception:
            pass


def mixed_patterns(text, N, m, seed, alphabet=None):
    """Patterns for parity tests: ~70% substrings of the text, the rest random (mostly absent),
    plus edge bytes (0x00, 0x01, 0xFF) sprinkled in."""
    rng = np.random.default_rng(seed)
    t = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    out = np.zeros((N, m), dtype=np.uint8)
    alpha = np.unique(t) if alphabet is None else np.asarray(alphabet, dtype=np.uint8)
    for i in range(N):
        u = rng.random()
        if u < 0.7 and t.size >= m:
            s = int(rng.integers(0, t.size - m + 1))
            out[i] = t[s:s + m]
            if u < 0.1 and m > 0:  # one mutated symbol
                out[i, int(rng.integers(0, m))] = alpha[int(rng.integers(0, alpha.size))]
        elif u < 0.95:
            out[i] = alpha[rng.integers(0, alpha.size, size=m)]
        else:
            out[i] = rng.integers(0, 256, size=m, dtype=np.uint8)
    return out.reshape(-1)


def repetitive_text(n, base, snps, seed, sigma=4):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGTNRYKMSWBDHV", dtype=np.uint8)[:sigma] if sigma <= 15 else np.arange(33, 33 + sigma, dtype=np.uint8)
    cur = alpha[rng.integers(0, alpha.size, size=base)]
    parts = []
    total = 0
    while total < n:
        parts.append(cur.copy())
        total += base
        for _ in range(snps):
            cur[int(rng.integers(0, base))] = alpha[int(rng.integers(0, alpha.size))]
    return np.concatenate(parts)[:n].copy()


needs_ref = pytest.mark.skipif(not os.path.exists(ob.REF_SO), reason="oracle/_ref not built (needs /root/reference)")
