"""CPU tier: the C-ABI library loads and exports every symbol include/rindex_gpu.h declares.
No compute calls here (they need a GPU); only calls that are defined without one."""
import ctypes
import os
import re

import pytest

from conftest import rib, ROOT


@pytest.fixture(scope="module")
def lib():
    rib.build_gpu()
    return ctypes.CDLL(rib.GPU_SO)


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ri[gh]_[a-z0-9_]+)\s*\(", src)))


def test_gpu_library_exports_every_declared_symbol(lib):
    names = _declared("rindex_gpu.h")
    assert set(names) == set(rib.DECLARED_SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None


def test_host_library_exports_every_declared_symbol():
    rib.build_host()
    hl = ctypes.CDLL(rib.HOST_SO)
    for n in _declared("rindex_host.h"):
        assert getattr(hl, n) is not None


def test_error_strings_and_no_device_behaviour(lib):
    lib.rig_strerror.restype = ctypes.c_char_p
    lib.rig_version.restype = ctypes.c_char_p
    assert lib.rig_strerror(0) == b"ok"
    assert lib.rig_strerror(-5) == b"logical arrays are not a valid r-index"
    assert b"sm_100a" in lib.rig_version()
    assert lib.rig_device_count() >= 0
    assert lib.rig_index_create(None, 0, None) == -1


def test_product_has_no_cpu_fallback_and_never_touches_the_oracle():
    """The product tree must not reference oracle/ (import, link, dlopen)."""
    pkg = os.path.join(ROOT, "r-index_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                assert "libri_oracle" not in s and "libri_ref" not in s and "oracle.bindings" not in s and "from oracle" not in s, f


def test_sass_is_sm100a(lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", rib.GPU_SO], capture_output=True, text=True).stdout
    assert "sm_100a" in out
