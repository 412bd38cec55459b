"""In-tree native builds (no JIT cache: the built .so files travel with the repo snapshot to the GPU box).

  libbvgraph_b200.so   CUDA kernels (sm_100a) + the C ABI of include/bvgraph_b200.h   [nvcc]
  libbvgraph_tools.so  host-side compressor + synthetic generator, include/bvgraph_tools.h [g++]
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")

CUDA_LIB = os.path.join(PKG, "libbvgraph_b200.so")
TOOLS_LIB = os.path.join(PKG, "libbvgraph_tools.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-Xptxas", "-v"]
# x86-64-v3 rather than -march=native: the .so is built here and runs on the GPU box's host CPU
GXX_FLAGS = ["-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(subdir, exts):
    d = os.path.join(CSRC, subdir)
    out = []
    for root, _, files in os.walk(d):
        out += [os.path.join(root, f) for f in sorted(files) if f.endswith(exts)]
    return out


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def tools_library(force=False):
    srcs = _sources("tools", (".cpp",))
    deps = srcs + [os.path.join(INCLUDE, "bvgraph_tools.h")]
    if force or _stale(TOOLS_LIB, deps):
        if not shutil.which("g++"):
            if os.path.exists(TOOLS_LIB):
                return TOOLS_LIB
            raise RuntimeError("g++ not found and libbvgraph_tools.so not prebuilt")
        subprocess.check_call(["g++"] + GXX_FLAGS + ["-o", TOOLS_LIB] + srcs)
    return TOOLS_LIB


def cuda_library(force=False, verbose=False):
    srcs = _sources("cuda", (".cu",))
    deps = srcs + _sources("cuda", (".cuh", ".hpp", ".h")) + [os.path.join(INCLUDE, "bvgraph_b200.h")]
    if force or _stale(CUDA_LIB, deps):
        nvcc = nvcc_path()
        if not os.path.exists(nvcc):
            if os.path.exists(CUDA_LIB):
                return CUDA_LIB
            raise RuntimeError("nvcc not found and libbvgraph_b200.so not prebuilt: the CUDA path is the only decode path")
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-o", CUDA_LIB] + srcs + ["-lcudart"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log = os.path.join(PKG, "csrc", "cuda", "ptxas.log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + res.stdout)
        if verbose or res.returncode:
            print(res.stdout)
        if res.returncode:
            raise RuntimeError("nvcc failed (see %s)" % log)
    return CUDA_LIB


def build_all(force=False, verbose=False):
    return tools_library(force), cuda_library(force, verbose)
