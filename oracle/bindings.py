"""ctypes bindings of the CHECKERS under oracle/ — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may
import this module. The product package (r-index_b200/) never does.

  RefIndex    oracle/_ref/libri_ref.so   the reference's own r_index<> code (compiled from
                                         /root/reference against oracle/sdsl_shim)   kind = "reference"
  PortIndex   oracle/libri_oracle.so     plain-C restatement (oracle/ri_oracle.c)     kind = "port"
"""
import ctypes
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libri_ref.so")
PORT_SO = os.path.join(HERE, "libri_oracle.so")

_u64 = ctypes.c_uint64
_vp = ctypes.c_void_p


def build(quiet=True):
    """Compile the checkers (make -f oracle/Makefile). _ref needs /root/reference; absent -> prebuilt kept."""
    out = subprocess.run(["make", "-f", os.path.join(HERE, "Makefile"), "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8
        return np.ascontiguousarray(buf)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def have_ref():
    return os.path.exists(REF_SO)


def have_port():
    return os.path.exists(PORT_SO)


_ref_lib = None


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        lib = ctypes.CDLL(REF_SO)
        lib.ref_build.restype = _vp
        lib.ref_build.argtypes = [_vp, _u64, ctypes.c_int]
        lib.ref_free.argtypes = [_vp]
        lib.ref_bwt_size.restype = _u64
        lib.ref_bwt_size.argtypes = [_vp]
        lib.ref_number_of_runs.restype = _u64
        lib.ref_number_of_runs.argtypes = [_vp]
        lib.ref_save.argtypes = [_vp, ctypes.c_char_p]
        lib.ref_load.restype = _vp
        lib.ref_load.argtypes = [ctypes.c_char_p]
        lib.ref_count_batch.restype = ctypes.c_double
        lib.ref_count_batch.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, ctypes.c_int]
        lib.ref_locate_batch.restype = ctypes.c_double
        lib.ref_locate_batch.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, ctypes.c_int]
        lib.ref_bwt_rank.restype = _u64
        lib.ref_bwt_rank.argtypes = [_vp, _u64, ctypes.c_uint8]
        lib.ref_bwt_at.restype = ctypes.c_uint8
        lib.ref_bwt_at.argtypes = [_vp, _u64]
        lib.ref_phi.restype = _u64
        lib.ref_phi.argtypes = [_vp, _u64]
        lib.ref_extract.argtypes = [_vp] * 7
        lib.ref_from_logical.restype = _vp
        lib.ref_from_logical.argtypes = [_u64, _u64, _vp, _vp, _vp, _vp, _vp]
        lib.ref_navigate.argtypes = [_vp, ctypes.c_int, _vp, _u64, _vp]
        lib.ref_get_bwt.argtypes = [_vp, _vp]
        lib.ref_break_range.restype = _u64
        lib.ref_break_range.argtypes = [_vp, _u64, _u64, ctypes.c_uint8, _vp, _vp]
        lib.ref_closest_run_break.restype = _u64
        lib.ref_closest_run_break.argtypes = [_vp, _u64, _u64, ctypes.c_uint8]
        _ref_lib = lib
    return _ref_lib


class RefIndex:
    """The reference's r_index<> (unmodified headers) behind a C driver."""

    kind = "reference"

    def __init__(self, handle):
        if not handle:
            raise ValueError("reference index could not be built/loaded")
        self.h = handle
        self.lib = ref_lib()
        self.n = self.lib.ref_bwt_size(self.h)
        self.r = self.lib.ref_number_of_runs(self.h)

    @classmethod
    def from_text(cls, text):
        t = _as_u8(text)
        return cls(ref_lib().ref_build(_ptr(t), t.size, 1))

    @classmethod
    def load(cls, path):
        return cls(ref_lib().ref_load(path.encode()))

    @classmethod
    def from_logical(cls, arrays):
        """The reference's structures (rle_string, sparse_sd_vector, int_vector: their own constructors) over a given
        BWT + samples, i.e. the r_index constructor without its suffix sort. arrays: dict as extract() returns."""
        a = {k: np.ascontiguousarray(arrays[k], dtype=(np.uint8 if k == "run_heads" else np.uint64))
             for k in ("run_heads", "run_lens", "samples_last", "pred_pos", "pred_to_run")}
        return cls(ref_lib().ref_from_logical(int(arrays["n"]), int(arrays["r"]), _ptr(a["run_heads"]), _ptr(a["run_lens"]),
                                              _ptr(a["samples_last"]), _ptr(a["pred_pos"]), _ptr(a["pred_to_run"])))

    def save(self, path):
        if self.lib.ref_save(self.h, path.encode()) != 0:
            raise IOError(path)

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count(self, patterns, N, m, threads=1, want=True):
        p = _as_u8(patterns)
        assert p.size >= N * m
        lo = np.zeros(N, dtype=np.uint64) if want else None
        hi = np.zeros(N, dtype=np.uint64) if want else None
        secs = self.lib.ref_count_batch(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), threads)
        return lo, hi, secs

    def locate(self, patterns, N, m, threads=1, want=True):
        """Returns (lo, hi, occ_offsets, occ, seconds_in_locate_loop)."""
        p = _as_u8(patterns)
        if not want:
            tot = _u64(0)
            secs = self.lib.ref_locate_batch(self.h, _ptr(p), N, m, None, None, ctypes.byref(tot), threads)
            return None, None, None, int(tot.value), secs
        lo, hi, _ = self.count(p, N, m, threads)
        nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
        off = np.zeros(N + 1, dtype=np.uint64)
        np.cumsum(nocc, out=off[1:])
        occ = np.zeros(int(off[-1]), dtype=np.uint64)
        tot = _u64(0)
        secs = self.lib.ref_locate_batch(self.h, _ptr(p), N, m, _ptr(off), _ptr(occ), ctypes.byref(tot), threads)
        assert int(tot.value) == int(off[-1])
        return lo, hi, off, occ, secs

    def rank(self, i, c):
        return self.lib.ref_bwt_rank(self.h, i, c)

    def bwt_at(self, i):
        return self.lib.ref_bwt_at(self.h, i)

    def phi(self, i):
        return self.lib.ref_phi(self.h, i)

    def navigate(self, op, positions):
        """op 0: operator[](i), 1: LF(i), 2: FL(i), 3: F_at(i) — the reference's own methods."""
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        out = np.empty(pos.size, dtype=np.uint64)
        self.lib.ref_navigate(self.h, op, _ptr(pos), pos.size, _ptr(out))
        return out

    def break_range(self, l, r, c):
        """rle_string::break_range through the reference's own method: list of (first, last)."""
        k = int(self.lib.ref_break_range(self.h, l, r, c, None, None))
        a = np.zeros(k, dtype=np.uint64); b = np.zeros(k, dtype=np.uint64)
        self.lib.ref_break_range(self.h, l, r, c, _ptr(a), _ptr(b))
        return a, b

    def closest_run_break(self, l, r, c):
        return int(self.lib.ref_closest_run_break(self.h, l, r, c))

    def get_bwt(self):
        out = np.empty(int(self.n), dtype=np.uint8)
        self.lib.ref_get_bwt(self.h, _ptr(out))
        return out

    def extract(self):
        r = int(self.r)
        F = np.zeros(257, dtype=np.uint64)
        heads = np.zeros(r, dtype=np.uint8)
        lens = np.zeros(r, dtype=np.uint64)
        sl = np.zeros(r, dtype=np.uint64)
        pp = np.zeros(r, dtype=np.uint64)
        ptr = np.zeros(r, dtype=np.uint64)
        self.lib.ref_extract(self.h, _ptr(F), _ptr(heads), _ptr(lens), _ptr(sl), _ptr(pp), _ptr(ptr))
        return dict(n=int(self.n), r=r, F=F, run_heads=heads, run_lens=lens, samples_last=sl, pred_pos=pp,
                    pred_to_run=ptr)


_port_lib = None


def port_lib():
    global _port_lib
    if _port_lib is None:
        lib = ctypes.CDLL(PORT_SO)
        lib.rio_build.restype = _vp
        lib.rio_build.argtypes = [_vp, _u64, _vp]
        lib.rio_free.argtypes = [_vp]
        lib.rio_n.restype = _u64
        lib.rio_n.argtypes = [_vp]
        lib.rio_r.restype = _u64
        lib.rio_r.argtypes = [_vp]
        lib.rio_count_batch.restype = ctypes.c_double
        lib.rio_count_batch.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp]
        lib.rio_locate_batch.restype = ctypes.c_double
        lib.rio_locate_batch.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp]
        lib.rio_extract.argtypes = [_vp] * 7
        lib.rio_check_sa.restype = ctypes.c_int
        lib.rio_check_sa.argtypes = [_vp, _u64, _vp]
        lib.rio_brute_count.restype = _u64
        lib.rio_brute_count.argtypes = [_vp, _u64, _vp, _u64]
        lib.rio_rank.restype = _u64
        lib.rio_rank.argtypes = [_vp, _u64, ctypes.c_uint8]
        lib.rio_phi.restype = _u64
        lib.rio_phi.argtypes = [_vp, _u64]
        _port_lib = lib
    return _port_lib


class PortIndex:
    """Plain-C restatement of the reference's count/locate path (oracle/ri_oracle.c)."""

    kind = "port"

    def __init__(self, text, sa=None):
        """sa: optional int64 suffix array of text+\\0 (validated with rio_check_sa before use);
        without it the oracle sorts suffixes itself (prefix doubling; fine up to a few MB)."""
        self.lib = port_lib()
        t = _as_u8(text)
        self._text = t
        if sa is not None:
            sa = np.ascontiguousarray(sa, dtype=np.int64)
            assert sa.size == t.size + 1
            if self.lib.rio_check_sa(_ptr(t), t.size, _ptr(sa)) != 0:
                raise ValueError("supplied suffix array failed verification")
        self.h = self.lib.rio_build(_ptr(t), t.size, _ptr(sa))
        if not self.h:
            raise ValueError("oracle build failed (reserved bytes 0x00/0x01?)")
        self.n = self.lib.rio_n(self.h)
        self.r = self.lib.rio_r(self.h)

    def close(self):
        if self.h:
            self.lib.rio_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count(self, patterns, N, m, threads=1, want=True):
        p = _as_u8(patterns)
        lo = np.zeros(N, dtype=np.uint64)
        hi = np.zeros(N, dtype=np.uint64)
        secs = self.lib.rio_count_batch(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi))
        return lo, hi, secs

    def locate(self, patterns, N, m, threads=1, want=True):
        p = _as_u8(patterns)
        lo, hi, _ = self.count(p, N, m)
        nocc = np.where(hi >= lo, hi - lo + np.uint64(1), np.uint64(0)).astype(np.uint64)
        off = np.zeros(N + 1, dtype=np.uint64)
        np.cumsum(nocc, out=off[1:])
        occ = np.zeros(int(off[-1]), dtype=np.uint64)
        secs = self.lib.rio_locate_batch(self.h, _ptr(p), N, m, _ptr(off), _ptr(occ), None)
        if not want:
            return None, None, None, int(occ.size), secs
        return lo, hi, off, occ, secs

    def rank(self, i, c):
        return self.lib.rio_rank(self.h, i, c)

    def phi(self, i):
        return self.lib.rio_phi(self.h, i)

    def extract(self):
        r = int(self.r)
        F = np.zeros(257, dtype=np.uint64)
        heads = np.zeros(r, dtype=np.uint8)
        lens = np.zeros(r, dtype=np.uint64)
        sl = np.zeros(r, dtype=np.uint64)
        pp = np.zeros(r, dtype=np.uint64)
        ptr = np.zeros(r, dtype=np.uint64)
        self.lib.rio_extract(self.h, _ptr(F), _ptr(heads), _ptr(lens), _ptr(sl), _ptr(pp), _ptr(ptr))
        return dict(n=int(self.n), r=r, F=F, run_heads=heads, run_lens=lens, samples_last=sl, pred_pos=pp,
                    pred_to_run=ptr)


def brute_count(text, pattern):
    """Overlapping occurrences of pattern in text by direct comparison (SDSL-independent witness;
    the idea of ri-locate -c, reference ri-locate.cpp:156-190)."""
    t = _as_u8(text)
    p = _as_u8(pattern)
    return int(port_lib().rio_brute_count(_ptr(t), t.size, _ptr(p), p.size))


def brute_locate_sorted(text, pattern):
    t = bytes(_as_u8(text))
    p = bytes(_as_u8(pattern))
    out = []
    i = t.find(p)
    while i >= 0:
        out.append(i)
        i = t.find(p, i + 1)
    return np.array(out, dtype=np.uint64)
