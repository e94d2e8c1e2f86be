"""ctypes binding of librindex_gpu.so (include/rindex_gpu.h) — the drop-in boundary.

No fallback: if the CUDA library is missing or a call fails, this raises. Nothing here
computes; it only marshals numpy host buffers / raw device pointers into the C ABI."""
import ctypes
import os
import numpy as np
from . import _build
from ._host import LogicalView, HostIndex

_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_vp = ctypes.c_void_p

# every symbol include/rindex_gpu.h declares (tests check the .so exports all of them)
DECLARED_SYMBOLS = [
    "rig_device_count", "rig_strerror", "rig_last_cuda_error", "rig_version", "rig_index_create",
    "rig_index_create_ex", "rig_index_destroy", "rig_index_info_get", "rig_count_batch", "rig_locate_batch",
    "rig_count_batch_dev", "rig_locate_batch_dev", "rig_digest_dev", "rig_last_timing", "rig_plan_batch_dev", "rig_expand_shard_dev",
    "rig_text_attach", "rig_sort_occurrences_dev", "rig_check_dev", "rig_locate_batch_ex",
    "rig_navigate_batch", "rig_navigate_batch_dev", "rig_get_bwt", "rig_locate_batch32",
    "rig_break_range_batch", "rig_closest_run_break_batch", "rig_fetch_occurrences", "rig_host_alloc", "rig_host_free",
    "rig_index_save_flat", "rig_index_load_flat", "rig_counts_dev", "rig_balanced_cuts_dev",
]

RIG_ERR_CAPACITY = -4


class Options(ctypes.Structure):
    _fields_ = [("runs_per_block", _u32), ("lf_bucket_log2", _u32), ("phi_bucket_log2", _u32),
                ("expand_threads", _u32), ("reserved", _u32 * 4)]


class IndexInfo(ctypes.Structure):
    _fields_ = [("n", _u64), ("r", _u64), ("sigma", _u64), ("device_bytes", _u64), ("lf_blocks", _u64),
                ("lf_buckets", _u64), ("phi_buckets", _u64), ("runs_per_block", _u32), ("lf_shift", _u32),
                ("phi_shift", _u32), ("device", _u32), ("sm_count", _u32), ("phi_jump", _u32),
                ("phi_jump_pieces", _u64), ("words32", _u32), ("lf_record_bytes", _u32), ("seed_jump", _u32),
                ("seed_shift", _u32), ("seed_pieces", _u64), ("seed_bytes", _u64)]


class Timing(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("search_ms", ctypes.c_float), ("scan_ms", ctypes.c_float),
                ("expand_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float), ("launches", _u32), ("slices", _u32),
                ("lf_steps", _u64), ("occ_total", _u64), ("chains", _u64), ("seed_ms", ctypes.c_float),
                ("window_ms", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class CheckReport(ctypes.Structure):
    """struct rig_check_report: the outcome of ri-locate's -c self-check run on the device."""
    _fields_ = [("patterns_checked", _u64), ("wrong_count_patterns", _u64), ("wrong_occurrences", _u64),
                ("unsorted_or_duplicate", _u64), ("first_bad_pattern", _u64), ("first_bad_position", _u64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}

    @property
    def clean(self):
        return self.wrong_count_patterns == 0 and self.wrong_occurrences == 0 and self.unsorted_or_duplicate == 0


LOCATE_SORT, LOCATE_CHECK, LOCATE_DEVICE_ONLY = 1, 2, 4
NAV_BWT, NAV_LF, NAV_FL, NAV_F_AT = 0, 1, 2, 3


class RigError(RuntimeError):
    def __init__(self, code, where):
        lib = gpu_lib()
        msg = lib.rig_strerror(code).decode()
        cuda = lib.rig_last_cuda_error().decode()
        super().__init__("%s: %s (%d)%s" % (where, msg, code, (" [" + cuda + "]") if cuda else ""))
        self.code = code


_lib = None


def gpu_lib():
    """Load librindex_gpu.so. Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.GPU_SO):
            raise RuntimeError("librindex_gpu.so is missing (run __graft_entry__.build()); "
                               "the query path has no CPU fallback")
        lib = ctypes.CDLL(_build.GPU_SO)
        lib.rig_device_count.restype = ctypes.c_int
        lib.rig_strerror.restype = ctypes.c_char_p
        lib.rig_strerror.argtypes = [ctypes.c_int]
        lib.rig_last_cuda_error.restype = ctypes.c_char_p
        lib.rig_version.restype = ctypes.c_char_p
        lib.rig_index_create.argtypes = [ctypes.POINTER(LogicalView), ctypes.c_int, ctypes.POINTER(_vp)]
        lib.rig_index_create_ex.argtypes = [ctypes.POINTER(LogicalView), ctypes.c_int, ctypes.POINTER(Options),
                                            ctypes.POINTER(_vp)]
        lib.rig_index_destroy.argtypes = [_vp]
        lib.rig_index_info_get.argtypes = [_vp, ctypes.POINTER(IndexInfo)]
        lib.rig_count_batch.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp]
        lib.rig_locate_batch.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64)]
        lib.rig_count_batch_dev.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp]
        lib.rig_locate_batch_dev.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64), _vp]
        lib.rig_digest_dev.argtypes = [_vp, _vp, _u64, ctypes.POINTER(_u64 * 2), _vp]
        lib.rig_last_timing.argtypes = [_vp, ctypes.POINTER(Timing)]
        lib.rig_plan_batch_dev.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _u32, _u64, _vp, ctypes.POINTER(_u64), _vp]
        lib.rig_plan_batch_dev.restype = ctypes.c_int
        lib.rig_expand_shard_dev.argtypes = [_vp, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64), _vp]
        lib.rig_expand_shard_dev.restype = ctypes.c_int
        lib.rig_navigate_batch.argtypes = [_vp, ctypes.c_int, _vp, _u64, _vp]
        lib.rig_navigate_batch_dev.argtypes = [_vp, ctypes.c_int, _vp, _u64, _vp, _vp]
        lib.rig_get_bwt.argtypes = [_vp, _u64, _u64, _vp]
        lib.rig_locate_batch32.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64), _u32]
        lib.rig_break_range_batch.argtypes = [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64)]
        lib.rig_closest_run_break_batch.argtypes = [_vp, _vp, _vp, _vp, _u64, _vp]
        lib.rig_index_save_flat.argtypes = [_vp, ctypes.c_char_p]
        lib.rig_index_load_flat.argtypes = [ctypes.c_char_p, ctypes.POINTER(LogicalView), ctypes.c_int, ctypes.POINTER(_vp)]
        lib.rig_counts_dev.argtypes = [_vp, _vp, _vp, _u64, _vp, _vp]
        lib.rig_balanced_cuts_dev.argtypes = [_vp, _vp, _u64, _u32, _u64, _vp, _vp]
        lib.rig_fetch_occurrences.argtypes = [_vp, _u64, _u64, _vp]
        lib.rig_host_alloc.restype = _vp
        lib.rig_host_alloc.argtypes = [_u64]
        lib.rig_host_free.argtypes = [_vp]
        lib.rig_text_attach.argtypes = [_vp, _vp, _u64]
        lib.rig_sort_occurrences_dev.argtypes = [_vp, _u64, _vp, _vp, _u64, _vp]
        lib.rig_check_dev.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64, ctypes.c_int,
                                      ctypes.POINTER(CheckReport), _vp]
        lib.rig_locate_batch_ex.argtypes = [_vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64), _u32,
                                            ctypes.POINTER(CheckReport)]
        _lib = lib
    return _lib


def device_count():
    return gpu_lib().rig_device_count()


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        assert buf.dtype == np.uint8
        return np.ascontiguousarray(buf)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def view_from_arrays(d):
    """rig_logical_view over a dict of numpy arrays (keys as HostIndex.arrays()). Returns (view, keepalive)."""
    keep = dict(
        F=np.ascontiguousarray(d["F"], dtype=np.uint64), run_heads=np.ascontiguousarray(d["run_heads"], dtype=np.uint8),
        run_lens=np.ascontiguousarray(d["run_lens"], dtype=np.uint64),
        samples_last=np.ascontiguousarray(d["samples_last"], dtype=np.uint64),
        pred_pos=np.ascontiguousarray(d["pred_pos"], dtype=np.uint64),
        pred_to_run=np.ascontiguousarray(d["pred_to_run"], dtype=np.uint64))
    v = LogicalView()
    v.n, v.r = int(d["n"]), int(d["r"])
    for k, a in keep.items():
        setattr(v, k, a.ctypes.data)
    return v, keep


class GpuIndex:
    """A flattened r-index resident in one GPU's HBM (struct rig_index)."""

    def __init__(self, source, device=0, runs_per_block=0, lf_bucket_log2=0, phi_bucket_log2=0, expand_threads=0,
                 phi_jump=0, seed_jump=0, flat=None):
        """flat: path of a file written by save_flat(); when it exists and belongs to `source` the flatten step is
        skipped (rig_index_load_flat), otherwise the index is flattened as usual."""
        lib = gpu_lib()
        if flat is not None and os.path.exists(flat):
            view = source.view if isinstance(source, HostIndex) else (view_from_arrays(source)[0] if source is not None else None)
            h = _vp()
            rc = lib.rig_index_load_flat(flat.encode(), ctypes.byref(view) if view is not None else None, device, ctypes.byref(h))
            if rc == 0:
                self.h, self.lib, self.device, self._keep = h, lib, device, None
                self.info = IndexInfo()
                lib.rig_index_info_get(self.h, ctypes.byref(self.info))
                self.n, self.r = int(self.info.n), int(self.info.r)
                self.from_flat = True
                return
            if rc != -5 or source is None:   # anything but "not this index's flat file" is an error
                raise RigError(rc, "rig_index_load_flat")
        self.from_flat = False
        if isinstance(source, HostIndex):
            view, self._keep = source.view, source
        elif isinstance(source, dict):
            view, self._keep = view_from_arrays(source)
        else:
            raise TypeError("GpuIndex needs a HostIndex or a dict of logical arrays")
        opt = Options(runs_per_block, lf_bucket_log2, phi_bucket_log2, expand_threads)
        opt.reserved[0] = phi_jump
        opt.reserved[2] = seed_jump
        h = _vp()
        rc = lib.rig_index_create_ex(ctypes.byref(view), device, ctypes.byref(opt), ctypes.byref(h))
        if rc != 0:
            raise RigError(rc, "rig_index_create_ex")
        self.h = h
        self.lib = lib
        self.device = device
        self._keep = None
        self.info = IndexInfo()
        lib.rig_index_info_get(self.h, ctypes.byref(self.info))
        self.n, self.r = int(self.info.n), int(self.info.r)

    def save_flat(self, path):
        rc = self.lib.rig_index_save_flat(self.h, path.encode())
        if rc != 0:
            raise RigError(rc, "rig_index_save_flat")

    def close(self):
        if getattr(self, "h", None):
            self.lib.rig_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer entry points (the reference-facing calls; copies are inside) ----
    def count(self, patterns, N, m, lo=None, hi=None):
        p = _as_u8(patterns)
        assert p.size >= N * m
        lo = np.empty(N, dtype=np.uint64) if lo is None else lo
        hi = np.empty(N, dtype=np.uint64) if hi is None else hi
        rc = self.lib.rig_count_batch(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi))
        if rc != 0:
            raise RigError(rc, "rig_count_batch")
        return lo, hi

    def locate(self, patterns, N, m, occ=None):
        """Returns (lo, hi, occ_offsets, occ). Two-call idiom when occ is not supplied."""
        p = _as_u8(patterns)
        assert p.size >= N * m
        lo = np.empty(N, dtype=np.uint64)
        hi = np.empty(N, dtype=np.uint64)
        off = np.empty(N + 1, dtype=np.uint64)
        tot = _u64(0)
        cap = 0 if occ is None else occ.size
        rc = self.lib.rig_locate_batch(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), _ptr(occ), cap,
                                       ctypes.byref(tot))
        if rc == RIG_ERR_CAPACITY:
            occ = np.empty(int(tot.value), dtype=np.uint64)
            rc = self.lib.rig_locate_batch(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), _ptr(occ), occ.size,
                                           ctypes.byref(tot))
        if rc != 0:
            raise RigError(rc, "rig_locate_batch")
        if occ is None:
            occ = np.empty(0, dtype=np.uint64)
        return lo, hi, off, occ[: int(tot.value)]

    # ---- single-position navigation (r_index::operator[], LF, FL, F_at; get_bwt) ----
    def navigate(self, op, positions):
        pos = np.ascontiguousarray(positions, dtype=np.uint64)
        out = np.empty(pos.size, dtype=np.uint64)
        rc = self.lib.rig_navigate_batch(self.h, op, _ptr(pos), pos.size, _ptr(out))
        if rc != 0:
            raise RigError(rc, "rig_navigate_batch")
        return out

    def break_range(self, lo, hi, c):
        """rle_string::break_range for a batch: returns (offsets[N+1], first[], last[])."""
        lo = np.ascontiguousarray(lo, dtype=np.uint64); hi = np.ascontiguousarray(hi, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint8)
        N = lo.size
        off = np.zeros(N + 1, dtype=np.uint64)
        tot = _u64(0)
        rc = self.lib.rig_break_range_batch(self.h, _ptr(lo), _ptr(hi), _ptr(c), N, _ptr(off), None, None, 0, ctypes.byref(tot))
        if rc not in (0, RIG_ERR_CAPACITY):
            raise RigError(rc, "rig_break_range_batch")
        first = np.zeros(int(tot.value), dtype=np.uint64); last = np.zeros(int(tot.value), dtype=np.uint64)
        if tot.value:
            rc = self.lib.rig_break_range_batch(self.h, _ptr(lo), _ptr(hi), _ptr(c), N, _ptr(off), _ptr(first), _ptr(last),
                                                first.size, ctypes.byref(tot))
            if rc != 0:
                raise RigError(rc, "rig_break_range_batch")
        return off, first, last

    def closest_run_break(self, lo, hi, c):
        lo = np.ascontiguousarray(lo, dtype=np.uint64); hi = np.ascontiguousarray(hi, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint8)
        out = np.zeros(lo.size, dtype=np.uint64)
        rc = self.lib.rig_closest_run_break_batch(self.h, _ptr(lo), _ptr(hi), _ptr(c), lo.size, _ptr(out))
        if rc != 0:
            raise RigError(rc, "rig_closest_run_break_batch")
        return out

    def get_bwt(self, start=0, length=None):
        length = self.n - start if length is None else length
        out = np.empty(length, dtype=np.uint8)
        rc = self.lib.rig_get_bwt(self.h, start, length, _ptr(out))
        if rc != 0:
            raise RigError(rc, "rig_get_bwt")
        return out

    # ---- ri-locate -o / -c post-processing on the device ----
    def text_attach(self, text):
        t = _as_u8(text)
        rc = self.lib.rig_text_attach(self.h, _ptr(t), t.size)
        if rc != 0:
            raise RigError(rc, "rig_text_attach")

    def locate_ex(self, patterns, N, m, flags):
        """rig_locate_batch_ex: returns (lo, hi, occ_offsets, occ, report-or-None)."""
        p = _as_u8(patterns)
        assert p.size >= N * m
        lo = np.empty(N, dtype=np.uint64)
        hi = np.empty(N, dtype=np.uint64)
        off = np.empty(N + 1, dtype=np.uint64)
        tot = _u64(0)
        rep = CheckReport()
        rc = self.lib.rig_locate_batch_ex(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), None, 0,
                                          ctypes.byref(tot), 0, None)
        if rc not in (0, RIG_ERR_CAPACITY):
            raise RigError(rc, "rig_locate_batch_ex")
        occ = np.empty(int(tot.value), dtype=np.uint64)
        rc = self.lib.rig_locate_batch_ex(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), _ptr(occ), occ.size,
                                          ctypes.byref(tot), flags, ctypes.byref(rep))
        if rc != 0:
            raise RigError(rc, "rig_locate_batch_ex")
        return lo, hi, off, occ, (rep if flags & LOCATE_CHECK else None)

    def locate_keep(self, patterns, N, m, flags=0):
        """rig_locate_batch_ex with RIG_LOCATE_DEVICE_ONLY: one call, occurrences stay on the device.
        Returns (lo, hi, occ_offsets, total, report-or-None); fetch positions with fetch()."""
        p = _as_u8(patterns)
        lo = np.empty(N, dtype=np.uint64); hi = np.empty(N, dtype=np.uint64); off = np.empty(N + 1, dtype=np.uint64)
        tot = _u64(0)
        rep = CheckReport()
        rc = self.lib.rig_locate_batch_ex(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), None, 0, ctypes.byref(tot),
                                          flags | LOCATE_DEVICE_ONLY, ctypes.byref(rep))
        if rc != 0:
            raise RigError(rc, "rig_locate_batch_ex")
        return lo, hi, off, int(tot.value), (rep if flags & LOCATE_CHECK else None)

    def fetch(self, first, count):
        out = np.empty(count, dtype=np.uint64)
        rc = self.lib.rig_fetch_occurrences(self.h, first, count, _ptr(out))
        if rc != 0:
            raise RigError(rc, "rig_fetch_occurrences")
        return out

    def sort_dev(self, N, d_off, d_occ, total, stream=None):
        rc = self.lib.rig_sort_occurrences_dev(self.h, N, d_off, d_occ, total, stream)
        if rc != 0:
            raise RigError(rc, "rig_sort_occurrences_dev")

    def check_dev(self, d_patterns, N, m, d_lo, d_hi, d_off, d_occ, total, is_sorted=True, stream=None):
        rep = CheckReport()
        rc = self.lib.rig_check_dev(self.h, d_patterns, N, m, d_lo, d_hi, d_off, d_occ, total, int(is_sorted),
                                    ctypes.byref(rep), stream)
        if rc != 0:
            raise RigError(rc, "rig_check_dev")
        return rep

    def locate32(self, patterns, N, m, flags=0):
        """rig_locate_batch32: (lo, hi, occ_offsets, occ as uint32)."""
        p = _as_u8(patterns)
        lo = np.empty(N, dtype=np.uint64)
        hi = np.empty(N, dtype=np.uint64)
        off = np.empty(N + 1, dtype=np.uint64)
        tot = _u64(0)
        rc = self.lib.rig_locate_batch32(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), None, 0, ctypes.byref(tot), 0)
        if rc not in (0, RIG_ERR_CAPACITY):
            raise RigError(rc, "rig_locate_batch32")
        occ = np.empty(int(tot.value), dtype=np.uint32)
        rc = self.lib.rig_locate_batch32(self.h, _ptr(p), N, m, _ptr(lo), _ptr(hi), _ptr(off), _ptr(occ), occ.size,
                                         ctypes.byref(tot), flags)
        if rc != 0:
            raise RigError(rc, "rig_locate_batch32")
        return lo, hi, off, occ

    def locate32_raw(self, p_ptr, N, m, lo_ptr, hi_ptr, off_ptr, occ_ptr, cap):
        tot = _u64(0)
        rc = self.lib.rig_locate_batch32(self.h, p_ptr, N, m, lo_ptr, hi_ptr, off_ptr, occ_ptr, cap, ctypes.byref(tot), 0)
        if rc != 0:
            raise RigError(rc, "rig_locate_batch32")
        return int(tot.value)

    def locate_raw(self, p_ptr, N, m, lo_ptr, hi_ptr, off_ptr, occ_ptr, cap):
        """Host-pointer call with caller-managed (e.g. pinned) buffers given as integer addresses."""
        tot = _u64(0)
        rc = self.lib.rig_locate_batch(self.h, p_ptr, N, m, lo_ptr, hi_ptr, off_ptr, occ_ptr, cap, ctypes.byref(tot))
        if rc != 0:
            raise RigError(rc, "rig_locate_batch")
        return int(tot.value)

    def count_raw(self, p_ptr, N, m, lo_ptr, hi_ptr):
        rc = self.lib.rig_count_batch(self.h, p_ptr, N, m, lo_ptr, hi_ptr)
        if rc != 0:
            raise RigError(rc, "rig_count_batch")

    # ---- device-buffer entry points (integer device addresses, e.g. torch .data_ptr()) ----
    def count_dev(self, d_patterns, N, m, d_lo, d_hi, stream=None):
        rc = self.lib.rig_count_batch_dev(self.h, d_patterns, N, m, d_lo, d_hi, stream)
        if rc != 0:
            raise RigError(rc, "rig_count_batch_dev")

    def locate_dev(self, d_patterns, N, m, d_lo, d_hi, d_off, d_occ, cap, stream=None):
        """Returns occ_total; raises RigError(code=-4) if cap is too small (needed count in .needed)."""
        tot = _u64(0)
        rc = self.lib.rig_locate_batch_dev(self.h, d_patterns, N, m, d_lo, d_hi, d_off, d_occ, cap, ctypes.byref(tot), stream)
        if rc != 0:
            e = RigError(rc, "rig_locate_batch_dev")
            e.needed = int(tot.value)
            raise e
        return int(tot.value)

    def counts_dev(self, d_lo, d_hi, N, d_nocc, stream=None):
        rc = self.lib.rig_counts_dev(self.h, d_lo, d_hi, N, d_nocc, stream)
        if rc != 0:
            raise RigError(rc, "rig_counts_dev")

    def balanced_cuts_dev(self, d_nocc, N, shards, cost=64, stream=None):
        cuts = (_u64 * (shards + 1))()
        rc = self.lib.rig_balanced_cuts_dev(self.h, d_nocc, N, shards, cost, cuts, stream)
        if rc != 0:
            raise RigError(rc, "rig_balanced_cuts_dev")
        return [int(x) for x in cuts]

    def digest_dev(self, d_values, count, stream=None):
        out = (_u64 * 2)()
        rc = self.lib.rig_digest_dev(self.h, d_values, count, ctypes.byref(out), stream)
        if rc != 0:
            raise RigError(rc, "rig_digest_dev")
        return int(out[0]), int(out[1])

    def plan_dev(self, d_patt, N, m, d_lo, d_hi, d_off, shards, cost=64, stream=None):
        """rig_plan_batch_dev on raw device pointers: (cuts list of shards + 1 entries, total occurrences of the batch)."""
        cuts = (_u64 * (shards + 1))()
        tot = _u64(0)
        rc = self.lib.rig_plan_batch_dev(self.h, d_patt, N, m, d_lo, d_hi, d_off, shards, cost, ctypes.cast(cuts, _vp), ctypes.byref(tot), stream)
        if rc != 0:
            raise RigError(rc, "rig_plan_batch_dev")
        return [int(c) for c in cuts], int(tot.value)

    def expand_shard_dev(self, N, c0, c1, d_lo, d_hi, d_off, d_occ, cap, stream=None):
        """rig_expand_shard_dev on raw device pointers: occurrences of the shard (RigError with .needed when cap is short)."""
        tot = _u64(0)
        rc = self.lib.rig_expand_shard_dev(self.h, N, c0, c1, d_lo, d_hi, d_off, d_occ, cap, ctypes.byref(tot), stream)
        if rc != 0:
            e = RigError(rc, "rig_expand_shard_dev")
            e.needed = int(tot.value)
            raise e
        return int(tot.value)

    def timing(self):
        t = Timing()
        self.lib.rig_last_timing(self.h, ctypes.byref(t))
        return t.as_dict()


def digest_host(values):
    """Host restatement of rig_digest_dev for checking: (sum, sum v*(i+1)) mod 2^64."""
    v = np.ascontiguousarray(values, dtype=np.uint64)
    idx = np.arange(1, v.size + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(v.sum(dtype=np.uint64)), int((v * idx).sum(dtype=np.uint64))
