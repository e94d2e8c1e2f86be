"""ctypes binding of librindex_host.so (include/rindex_host.h): index construction, container
I/O and the synthetic workloads of SURVEY.md §8d. Host-only; no query logic here."""
import ctypes
import os
import numpy as np
from . import _build

_u64 = ctypes.c_uint64
_vp = ctypes.c_void_p


class LogicalView(ctypes.Structure):
    """struct rig_logical_view (include/rindex_gpu.h)."""
    _fields_ = [("n", _u64), ("r", _u64), ("F", _vp), ("run_heads", _vp), ("run_lens", _vp),
                ("samples_last", _vp), ("pred_pos", _vp), ("pred_to_run", _vp)]


_lib = None


def host_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.HOST_SO):
            _build.build_host()
        lib = ctypes.CDLL(_build.HOST_SO)
        lib.rih_build_from_text.argtypes = [_vp, _u64, ctypes.POINTER(_vp)]
        lib.rih_build_from_text_pfp.argtypes = [_vp, _u64, ctypes.c_uint32, ctypes.c_uint32, _vp, ctypes.POINTER(_vp)]
        lib.rih_build_auto.argtypes = [_vp, _u64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_vp)]
        lib.rih_destroy.argtypes = [_vp]
        lib.rih_view.argtypes = [_vp, ctypes.POINTER(LogicalView)]
        lib.rih_save.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int]
        lib.rih_load.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_vp)]
        lib.rih_gen_text.argtypes = [ctypes.c_int, _u64, _u64, _u64, _u64, _vp]
        lib.rih_gen_patterns.argtypes = [_vp, _u64, _u64, _u64, _u64, _u64, _vp]
        lib.rih_suffix_array.argtypes = [_vp, _u64, _vp]
        _lib = lib
    return _lib


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8
        return np.ascontiguousarray(buf)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _np_view(addr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.uint64: ctypes.c_uint64, np.uint8: ctypes.c_uint8}[dtype]
    return np.ctypeslib.as_array(ctypes.cast(addr, ctypes.POINTER(ct)), shape=(count,))


class _Arrays(dict):
    """dict of borrowed numpy views + a reference to the owning HostIndex."""


class HostIndex:
    """Logical r-index on the host (what the reference keeps in F / bwt / pred / samples_last /
    pred_to_run, internal/r_index.hpp:655-665), built by this repo's own builder."""

    def __init__(self, handle):
        self.h = handle
        self.view = LogicalView()
        rc = host_lib().rih_view(self.h, ctypes.byref(self.view))
        if rc != 0:
            raise RuntimeError("rih_view failed: %d" % rc)
        self.n = int(self.view.n)
        self.r = int(self.view.r)

    @classmethod
    def from_text(cls, text):
        t = _as_u8(text)
        h = _vp()
        rc = host_lib().rih_build_from_text(_ptr(t), t.size, ctypes.byref(h))
        if rc == -1:
            raise ValueError("input string contains one of the reserved characters 0x0, 0x1")
        if rc != 0:
            raise RuntimeError("rih_build_from_text failed: %d" % rc)
        return cls(h)

    @classmethod
    def from_text_auto(cls, text):
        """The builder ri-build uses: prefix-free parsing for large texts, SA-IS otherwise (.used_pfp tells which)."""
        t = _as_u8(text)
        h = _vp()
        used = ctypes.c_int(0)
        rc = host_lib().rih_build_auto(_ptr(t), t.size, ctypes.byref(used), ctypes.byref(h))
        if rc == -1:
            raise ValueError("input string contains one of the reserved characters 0x0, 0x1")
        if rc != 0:
            raise RuntimeError("rih_build_auto failed: %d" % rc)
        obj = cls(h)
        obj.used_pfp = bool(used.value)
        return obj

    @classmethod
    def from_text_pfp(cls, text, w=0, p=0):
        """Prefix-free-parsing builder (same arrays as from_text); .pfp_stats holds the parse statistics."""
        t = _as_u8(text)
        h = _vp()
        stats = np.zeros(6, dtype=np.uint64)
        rc = host_lib().rih_build_from_text_pfp(_ptr(t), t.size, w, p, _ptr(stats), ctypes.byref(h))
        if rc == -1:
            raise ValueError("input string contains one of the reserved characters 0x0, 0x1")
        if rc != 0:
            raise RuntimeError("rih_build_from_text_pfp failed: %d" % rc)
        obj = cls(h)
        obj.pfp_stats = dict(zip(("phrases", "dict_bytes", "parse_len", "groups", "uniform_rows", "merged_rows"),
                                 (int(x) for x in stats)))
        return obj

    @classmethod
    def load(cls, path, with_flag_byte=True):
        h = _vp()
        rc = host_lib().rih_load(path.encode(), int(with_flag_byte), ctypes.byref(h))
        if rc != 0:
            raise IOError("rih_load(%s) failed: %d" % (path, rc))
        return cls(h)

    def save(self, path, with_flag_byte=True):
        rc = host_lib().rih_save(self.h, path.encode(), int(with_flag_byte))
        if rc != 0:
            raise IOError("rih_save(%s) failed: %d" % (path, rc))

    def arrays(self):
        """Borrowed numpy views of the logical arrays (valid while this object lives)."""
        v, r = self.view, self.r
        d = _Arrays(n=self.n, r=r, F=_np_view(v.F, 257, np.uint64), run_heads=_np_view(v.run_heads, r, np.uint8),
                    run_lens=_np_view(v.run_lens, r, np.uint64), samples_last=_np_view(v.samples_last, r, np.uint64),
                    pred_pos=_np_view(v.pred_pos, r, np.uint64), pred_to_run=_np_view(v.pred_to_run, r, np.uint64))
        d._owner = self  # the views borrow the C++ object's memory: the dict keeps it alive
        return d

    def close(self):
        if self.h:
            host_lib().rih_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


KINDS = {"dna_drift": 0, "dna_indep": 1, "versioned_doc": 2, "pangenome": 3}


def gen_text(kind, n, p0, p1, seed):
    out = np.zeros(n, dtype=np.uint8)
    rc = host_lib().rih_gen_text(KINDS[kind], n, p0, p1, seed, _ptr(out))
    if rc != 0:
        raise ValueError("rih_gen_text failed: %d" % rc)
    return out


def gen_patterns(text, N, m, seed, start_limit=0):
    t = _as_u8(text)
    out = np.zeros(N * m, dtype=np.uint8)
    rc = host_lib().rih_gen_patterns(_ptr(t), t.size, N, m, start_limit, seed, _ptr(out))
    if rc != 0:
        raise ValueError("rih_gen_patterns failed: %d" % rc)
    return out


def suffix_array(text):
    t = _as_u8(text)
    sa = np.zeros(t.size + 1, dtype=np.int64)
    rc = host_lib().rih_suffix_array(_ptr(t), t.size, _ptr(sa))
    if rc != 0:
        raise ValueError("rih_suffix_array failed: %d" % rc)
    return sa


def parse_pattern_file(data):
    """Pizza&Chili pattern file -> (N, m, bodies). Mirrors the reference's header parser
    (internal/utils.hpp:57-91: text after 'number=' / 'length=' up to the next space, atoi)."""
    data = bytes(data)
    nl = data.index(b"\n")
    header = data[:nl].decode("latin-1")

    def field(key):
        s = header.find(key)
        if s < 0 or s + len(key) >= len(header):
            raise ValueError("malformed header in patterns file")
        rest = header[s + len(key):]
        e = rest.find(" ")
        if e < 0:
            raise ValueError("malformed header in patterns file")
        digits = ""
        for ch in rest[:e].lstrip():
            if ch.isdigit() or (not digits and ch in "+-"):
                digits += ch
            else:
                break
        try:
            return int(digits)
        except ValueError:
            return 0

    N, m = field("number="), field("length=")
    body = np.frombuffer(data[nl + 1: nl + 1 + N * m], dtype=np.uint8)
    return N, m, body


def write_pattern_file(path, bodies, N, m, label="synthetic"):
    with open(path, "wb") as f:
        f.write(("# number=%d length=%d file=%s forbidden=\n" % (N, m, label)).encode())
        f.write(bytes(_as_u8(bodies)[: N * m]))
