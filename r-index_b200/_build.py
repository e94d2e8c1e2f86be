"""Compile the native libraries in-tree (so the built .so files travel with the gpurun snapshot).

  librindex_gpu.so   csrc/rindex_gpu.cu   nvcc -gencode arch=compute_100a,code=sm_100a  (the product)
  librindex_host.so  host/rindex_host.cpp g++                                           (builder/generators)
  cli/ri-build, ri-count, ri-locate       host C++ mains linked against both
"""
import os
import platform
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
INC = os.path.join(ROOT, "include")
GPU_SO = os.path.join(PKG, "librindex_gpu.so")
HOST_SO = os.path.join(PKG, "librindex_host.so")
CLI_DIR = os.path.join(PKG, "bin")

# host code: AVX2-class x86 (built here, run on the GPU box: not -march=native); other hosts (aarch64 Grace) get the default
MARCH = ["-march=x86-64-v3"] if platform.machine() in ("x86_64", "AMD64") else []

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd, log=None):
    out = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), out.stdout, out.stderr))
    return out.stdout + out.stderr


def _find(tool, fallback):
    p = shutil.which(tool)
    return p if p else fallback


def build_host(force=False):
    srcs = [os.path.join(PKG, "host", f) for f in ("rindex_host.cpp", "logical_index.hpp", "sais.hpp", "textgen.hpp", "pfp_builder.hpp")]
    srcs += [os.path.join(INC, "rindex_host.h"), os.path.join(INC, "rindex_gpu.h")]
    if force or _newer(HOST_SO, srcs):
        _run(["/usr/bin/g++", "-O3"] + MARCH + ["-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I", INC,
              "-o", HOST_SO, srcs[0]])
    return HOST_SO


def build_gpu(force=False):
    csrc = os.path.join(PKG, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh", ".hpp", ".h"))]
    srcs.append(os.path.join(INC, "rindex_gpu.h"))
    if force or _newer(GPU_SO, srcs):
        nvcc = _find("nvcc", "/usr/local/cuda/bin/nvcc")
        _run([nvcc] + NVCC_FLAGS + ["-shared", "-ccbin", "/usr/bin/g++", "-I", INC, "-o", GPU_SO,
                                     os.path.join(csrc, "rindex_gpu.cu"), "-lcudart", "-Xcompiler", "-pthread"],
             log=os.path.join(PKG, "csrc", "ptxas.log"))
    return GPU_SO


def build_cli(force=False):
    os.makedirs(CLI_DIR, exist_ok=True)
    hdrs = [os.path.join(PKG, "host", f) for f in os.listdir(os.path.join(PKG, "host"))]
    outs = []
    for name in ("ri-build", "ri-count", "ri-locate", "ri-gen"):
        src = os.path.join(PKG, "cli", name + ".cpp")
        if not os.path.exists(src):
            continue
        out = os.path.join(CLI_DIR, name)
        if force or _newer(out, [src, GPU_SO] + hdrs):
            _run(["/usr/bin/g++", "-O3"] + MARCH + ["-std=c++17", "-w", "-pthread", "-I", INC, "-I", os.path.join(PKG, "host"),
                  "-o", out, src, "-L", PKG, "-lrindex_gpu", "-Wl,-rpath,$ORIGIN/..", "-L/usr/local/cuda/lib64"])
        outs.append(out)
    return outs


def build_all(force=False):
    build_host(force)
    build_gpu(force)
    build_cli(force)
