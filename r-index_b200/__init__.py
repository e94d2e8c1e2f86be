"""rindex_b200 — B200-native batch count + locate engine for the r-index hot path.

The product is C++/CUDA: `librindex_gpu.so` (hand-written sm_100a kernels behind the C ABI of
include/rindex_gpu.h) plus the C++ host mirror of ri::r_index<> and the ri-build / ri-count /
ri-locate tools (r-index_b200/host, r-index_b200/cli). This Python package is the thin ctypes
harness tests/, bench.py and __graft_entry__.py drive it through; it holds no query logic and
no CPU fallback: every count/locate call goes to the CUDA library or raises.

The directory name has a hyphen (it mirrors the reference repo's name), so import it through
`__graft_entry__.load_package()` (registers it as module `rindex_b200`).
"""
from ._build import build_all, build_gpu, build_host, build_cli, GPU_SO, HOST_SO, CLI_DIR  # noqa: F401
from ._host import HostIndex, gen_text, gen_patterns, suffix_array, parse_pattern_file, write_pattern_file  # noqa: F401
from ._gpu import (GpuIndex, RigError, CheckReport, LOCATE_SORT, LOCATE_CHECK, LOCATE_DEVICE_ONLY, NAV_BWT, NAV_LF, NAV_FL, NAV_F_AT,  # noqa: F401
                   device_count, gpu_lib, DECLARED_SYMBOLS)

__all__ = ["HostIndex", "GpuIndex", "RigError", "gen_text", "gen_patterns", "suffix_array", "device_count",
           "build_all", "parse_pattern_file", "write_pattern_file"]
