"""Shared test plumbing.  `-m gpu` tests need a B200; everything else runs on CPU.

The oracle (oracle/: numpy restatement + oracle/_ref, the UNMODIFIED reference compiled from
/root/reference by oracle/Makefile) is test infrastructure: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline/reference arm may use it.
"""
import importlib.util
import json
import os
import subprocess
import sys
import hashlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEODIR = os.path.join(ROOT, "tests", "golden", "geometries")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def load_pyscuff():
    spec = importlib.util.spec_from_file_location("pyscuff", os.path.join(ROOT, "scuff-em_b200", "pyscuff.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ps():
    return load_pyscuff()


@pytest.fixture(scope="session")
def oracle():
    spec = importlib.util.spec_from_file_location("scuff_oracle", os.path.join(ROOT, "oracle", "scuff_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def have_ref():
    return os.path.exists(REF_DRIVER)


def ref_matrix(geo, omega, kbloch=None, imag=False, block=None, cache_in=None, cache_out=None, tmpdir="/tmp/scuff_ref_cache", env=None, trans=None):
    """Run the UNMODIFIED reference (oracle/_ref/ref_driver) and return (M, info).  Results are cached on disk."""
    if not have_ref():
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    os.makedirs(tmpdir, exist_ok=True)
    key = hashlib.md5(repr((geo, complex(omega), kbloch, imag, block, cache_in, sorted((env or {}).items()), trans)).encode()).hexdigest()[:16]
    out = os.path.join(tmpdir, f"{os.path.basename(geo)}.{key}.bin")
    meta = out + ".json"
    if not (os.path.exists(out) and os.path.exists(meta)) or cache_out:
        cmd = [REF_DRIVER, "matrix", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), out]
        if kbloch is not None:
            cmd += ["--kbloch", repr(float(kbloch[0])), repr(float(kbloch[1]))]
        if imag:
            cmd += ["--imag"]
        if block is not None:
            cmd += ["--block", str(block[0]), str(block[1])]
        if cache_in:
            cmd += ["--cache-in", cache_in]
        if cache_out:
            cmd += ["--cache-out", cache_out]
        if trans is not None:
            cmd += ["--transfile", trans[0], "--trans", str(trans[1])]
        r = subprocess.run(cmd, cwd=os.path.dirname(os.path.abspath(geo)), capture_output=True, text=True, check=True,
                           env=dict(os.environ, **(env or {})))
        open(meta, "w").write(r.stdout.strip().splitlines()[-1])
    info = json.load(open(meta))
    dt = np.float64 if info["real"] else np.complex128
    M = np.fromfile(out, dtype=dt).reshape(info["NC"], info["NR"]).T
    return M, info


def ref_ssi(geo, omega, nsa, nsb, L, cache_out=None, tmpdir="/tmp/scuff_ref_cache", env=None):
    """Reference GetSurfaceSurfaceInteractions for one surface pair with surface b displaced by L (ref_driver ssi)."""
    if not have_ref():
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    os.makedirs(tmpdir, exist_ok=True)
    key = hashlib.md5(repr(("ssi", geo, complex(omega), nsa, nsb, tuple(L))).encode()).hexdigest()[:16]
    out = os.path.join(tmpdir, f"{os.path.basename(geo)}.{key}.bin")
    cmd = [REF_DRIVER, "ssi", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), str(nsa), str(nsb)]
    cmd += [repr(float(x)) for x in L] + [out]
    if cache_out:
        cmd += ["--cache-out", cache_out]
    r = subprocess.run(cmd, cwd=os.path.dirname(os.path.abspath(geo)), capture_output=True, text=True, check=True,
                       env=dict(os.environ, **(env or {})))
    info = json.loads(r.stdout.strip().splitlines()[-1])
    return np.fromfile(out, dtype=np.complex128).reshape(info["NC"], info["NR"]).T, info


def ref_gradblock(geo, omega, nsa, nsb, imag=False, torque_axes=(), trans=None, tmpdir="/tmp/scuff_ref_cache", env=None):
    """Reference AssembleBEMMatrixBlock with GradM[3] and dMdT (ref_driver gradblock): returns [M, dM/dx, dM/dy, dM/dz, dM/dT_0, ...]"""
    if not have_ref():
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    os.makedirs(tmpdir, exist_ok=True)
    key = hashlib.md5(repr(("grad", geo, complex(omega), nsa, nsb, imag, tuple(map(tuple, torque_axes)), trans)).encode()).hexdigest()[:16]
    out = os.path.join(tmpdir, f"{os.path.basename(geo)}.{key}.bin")
    cmd = [REF_DRIVER, "gradblock", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), str(nsa), str(nsb), out]
    if imag:
        cmd += ["--imag"]
    if trans is not None:
        cmd += ["--transfile", trans[0], "--trans", str(trans[1])]
    for ax in torque_axes:
        cmd += ["--torque"] + [repr(float(x)) for x in ax]
    r = subprocess.run(cmd, cwd=os.path.dirname(os.path.abspath(geo)), capture_output=True, text=True, check=True,
                       env=dict(os.environ, **(env or {})))
    info = json.loads(r.stdout.strip().splitlines()[-1])
    dt = np.float64 if info["real"] else np.complex128
    A = np.fromfile(out, dtype=dt).reshape(4 + info["ntorque"], info["NC"], info["NR"])
    return [A[q].T for q in range(A.shape[0])], info


def ref_pbcinner(geo, omega, kbloch, nsa, nsb, cache_out=None, tmpdir="/tmp/scuff_ref_cache", env=None):
    """Reference AssembleBEMMatrixBlock of a periodic geometry minus its outer-cell (GBar) step, both from the reference's
    own code (ref_driver pbcinner): the Bloch-phased inner-cell blocks.  Returns (inner, full, info)."""
    if not have_ref():
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    os.makedirs(tmpdir, exist_ok=True)
    key = hashlib.md5(repr(("pbcinner", geo, complex(omega), tuple(kbloch), nsa, nsb)).encode()).hexdigest()[:16]
    out = os.path.join(tmpdir, f"{os.path.basename(geo)}.{key}.bin")
    cmd = [REF_DRIVER, "pbcinner", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))),
           repr(float(kbloch[0])), repr(float(kbloch[1])), str(nsa), str(nsb), out]
    if cache_out:
        cmd += ["--cache-out", cache_out]
    r = subprocess.run(cmd, cwd=os.path.dirname(os.path.abspath(geo)), capture_output=True, text=True, check=True,
                       env=dict(os.environ, **(env or {})))
    info = json.loads(r.stdout.strip().splitlines()[-1])
    rd = lambda f: np.fromfile(f, dtype=np.complex128).reshape(info["NC"], info["NR"]).T
    return rd(out), rd(out + ".full"), info


def rel_fro(A, B):
    return np.linalg.norm(A - B) / np.linalg.norm(B)


def ref_scatter(geo, omega, pw, ps_src=None, interior=False, method=0, trans=None, nopft=False, tmpdir="/tmp/scuff_ref_cache", env=None,
                points=None, rf=False, dsi=None, kbloch=None, pw_region=None):
    """One frequency point of the reference's scuff-scatter flow (ref_driver scatter): AssembleBEMMatrix, AssembleRHSVector,
    LU solve, GetPFTMatrix (EMT method).  pw = (E0[3] complex, nHat[3]); ps_src = (X0[3], P[3] complex) or None.
    Returns (RHS, KN, PFT[NS, 8], info).  Cached on disk."""
    if not have_ref():
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    os.makedirs(tmpdir, exist_ok=True)
    pkey = None if points is None else hashlib.md5(np.ascontiguousarray(points, dtype=np.float64).tobytes()).hexdigest()
    key = hashlib.md5(repr(("scatter", geo, complex(omega), pw, ps_src, interior, method, trans, nopft, pkey, rf, dsi, kbloch, pw_region)).encode()).hexdigest()[:16]
    out = os.path.join(tmpdir, f"{os.path.basename(geo)}.{key}.bin")
    meta = out + ".json"
    if not (os.path.exists(out) and os.path.exists(meta)):
        cmd = [REF_DRIVER, "scatter", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), out, "--pw"]
        for z in pw[0]:
            cmd += [repr(float(np.real(z))), repr(float(np.imag(z)))]
        cmd += [repr(float(x)) for x in pw[1]]
        if pw_region is not None:
            cmd += ["--region", pw_region]
        if ps_src is not None:
            cmd += ["--ps"] + [repr(float(x)) for x in ps_src[0]]
            for z in ps_src[1]:
                cmd += [repr(float(np.real(z))), repr(float(np.imag(z)))]
        if interior:
            cmd += ["--interior"]
        if nopft:
            cmd += ["--nopft"]
        cmd += ["--method", str(method)]
        if trans is not None:
            cmd += ["--transfile", trans[0], "--trans", str(trans[1])]
        if dsi is not None:
            cmd += ["--dsi", repr(float(dsi[0])), str(int(dsi[1]))]
        if kbloch is not None:
            cmd += ["--kbloch", repr(float(kbloch[0])), repr(float(kbloch[1]))]
        if points is not None:
            np.ascontiguousarray(points, dtype=np.float64).tofile(out + ".pts")
            cmd += ["--fields", out + ".pts"] + (["--rf"] if rf else [])
        r = subprocess.run(cmd, cwd=os.path.dirname(os.path.abspath(geo)), capture_output=True, text=True, check=True,
                           env=dict(os.environ, **(env or {})))
        open(meta, "w").write(r.stdout.strip().splitlines()[-1])
    info = json.load(open(meta))
    N, NS = info["N"], info["NS"]
    raw = np.fromfile(out, dtype=np.float64)
    RHS = raw[:2 * N].view(np.complex128).copy()
    KN = raw[2 * N:4 * N].view(np.complex128).copy()
    PFT = raw[4 * N:4 * N + 8 * NS].reshape(NS, 8).copy()
    o = 4 * N + 8 * NS
    if dsi is not None:
        info["dsipft"] = raw[o:o + 8 * NS].reshape(NS, 8).copy()
        o += 8 * NS
    if points is not None:
        NX = np.asarray(points).reshape(-1, 3).shape[0]
        info["fields"] = raw[o:o + 12 * NX].view(np.complex128).reshape(NX, 6).copy()
        if rf:
            o += 12 * NX
            info["rf"] = raw[o:o + 2 * N * 6 * NX].view(np.complex128).reshape(6 * NX, N).T.copy()
    return RHS, KN, PFT, info


def ref_dgf(geo, omega, X, total=False, tmpdir="/tmp/scuff_ref_cache", env=None):
    """Reference RWGGeometry::GetDyadicGFs (ref_driver dgf) at the rows of X (NX x 3 or NX x 6): returns (GE, GM), each (NX, 3, 3)."""
    if not have_ref():
        pytest.skip("oracle/_ref not built (run make -C oracle where /root/reference exists)")
    os.makedirs(tmpdir, exist_ok=True)
    X = np.ascontiguousarray(X, dtype=np.float64)
    key = hashlib.md5(repr(("dgf", geo, complex(omega), hashlib.md5(X.tobytes()).hexdigest(), X.shape, total)).encode()).hexdigest()[:16]
    out = os.path.join(tmpdir, f"{os.path.basename(geo)}.{key}.bin")
    if not os.path.exists(out):
        X.tofile(out + ".pts")
        cmd = [REF_DRIVER, "dgf", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), out + ".pts", str(X.shape[1]), out]
        if total:
            cmd.append("--total")
        subprocess.run(cmd, cwd=os.path.dirname(os.path.abspath(geo)), capture_output=True, text=True, check=True, env=dict(os.environ, **(env or {})))
    A = np.fromfile(out, dtype=np.complex128).reshape(-1, 18)
    return A[:, :9].reshape(-1, 3, 3), A[:, 9:].reshape(-1, 3, 3)
