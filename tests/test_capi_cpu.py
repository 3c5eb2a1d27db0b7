"""The C-ABI shared library loads and exports every symbol include/libscuff_cuda.h declares; without a CUDA device
the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import pytest
from conftest import ROOT


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scuffcuda_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(ps):
    L = ps.cuda_lib()
    syms = declared_symbols("libscuff_cuda.h")
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libscuff_cuda.so does not export " + s


def test_host_facade_loads(ps):
    H = ps.host_lib()
    for s in ("scuffb200_geometry_open", "scuffb200_assemble", "scuffb200_assemble_block", "scuffb200_lu_factorize",
              "scuffb200_preload_cache", "scuffb200_store_cache"):
        assert hasattr(H, s)


def test_no_device_no_fallback(ps):
    L = ps.cuda_lib()
    if L.scuffcuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert L.scuffcuda_create(0, C.byref(ctx)) != 0
    assert b"no CPU fallback" in L.scuffcuda_last_error(None)
    # the reference-facing API raises instead of computing anything on the host
    G = ps.RWGGeometry(os.path.join(ROOT, "tests", "golden", "geometries", "PECSphere_255.scuffgeo"))
    with pytest.raises(ps.ScuffError):
        G.AssembleBEMMatrix(1.0)


def test_ssi_args_layout(ps):
    """ctypes mirror of scuffcuda_ssi_args has the size the C compiler gives the struct"""
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "libscuff_cuda.h"\nint main(){printf("%zu", sizeof(scuffcuda_ssi_args));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        n = int(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert n == C.sizeof(ps.SSIArgs)


def test_argument_checks_before_any_device_work(ps):
    """errors the reference raises from AssembleBEMMatrix itself (AssembleBEMMatrix.cc:692-695) need no device"""
    G = ps.RWGGeometry(os.path.join(ROOT, "tests", "golden", "geometries", "PECSphere_255.scuffgeo"))
    with pytest.raises(ps.ScuffError, match="Bloch wavevector is undefined for compact"):
        G.AssembleBEMMatrix(1.0, kBloch=(0.1, 0.2))
    P = ps.RWGGeometry(os.path.join(ROOT, "tests", "golden", "geometries", "SiSlab_40.scuffgeo"))
    assert P.LDim == 2
    with pytest.raises(ps.ScuffError, match="Bloch wavevector must be specified"):
        P.AssembleBEMMatrix(1.0)
