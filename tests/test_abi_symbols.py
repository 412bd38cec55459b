"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/bvgraph_b200.h
declares, and fails loudly (BVG_ECUDA) instead of falling back when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import CNR, ROOT
from webgraph_b200 import bvgraph, build


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bvgt?_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    lib = C.CDLL(build.cuda_library())
    names = _declared("bvgraph_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(bvgraph.SYMBOLS)


def test_tools_library_exports_every_declared_symbol():
    lib = C.CDLL(build.tools_library())
    for n in _declared("bvgraph_tools.h"):
        assert hasattr(lib, n), n


def test_cuda_library_is_sm100a_only():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "--list-elf", build.cuda_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_strerror_and_launch_counter():
    lib = bvgraph.lib()
    assert b"IllegalArgumentException" in lib.bvg_strerror(-1)
    assert b"no CPU decode path" in lib.bvg_strerror(-7) or b"CUDA" in lib.bvg_strerror(-7)
    assert lib.bvg_kernel_launches() >= 0


def test_bad_properties_are_rejected_before_cuda(tmp_path):
    base = str(tmp_path / "x")
    open(base + ".properties", "w").write("graphclass=it.unimi.dsi.webgraph.EFGraph\nversion=0\n")
    with pytest.raises(bvgraph.FormatError):
        bvgraph.BVGraph.load(base)  # BVGraph.java:1528
    open(base + ".properties", "w").write("graphclass=it.unimi.dsi.webgraph.BVGraph\nnodes=1\narcs=0\nwindowsize=7\nmaxrefcount=3\nminintervallength=4\n")
    with pytest.raises(bvgraph.FormatError):
        bvgraph.BVGraph.load(base)  # missing version, :1533
    open(base + ".properties", "w").write("graphclass=it.unimi.dsi.webgraph.BVGraph\nversion=0\nnodes=1\narcs=0\nwindowsize=7\nmaxrefcount=3\n"
                                          "minintervallength=4\ncompressionflags=RESIDUALS_FOO\n")
    with pytest.raises(bvgraph.FormatError):
        bvgraph.BVGraph.load(base)  # :1362
    with pytest.raises(IOError):
        bvgraph.BVGraph.load(str(tmp_path / "missing"))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bvgraph.CudaError):
        bvgraph.BVGraph.load(CNR)
