"""The drop-in boundary, for real: the REFERENCE's own libscuff compiled with the binding of INTEGRATION.md section B
(integration/AssembleBEMMatrixCUDA.cc + a hook in RWGGeometry::AssembleBEMMatrixBlock; oracle/Makefile target `cuda`) runs its
block assembly through libscuff_cuda.so when SCUFF_CUDA=1.  Checked here on the B200:
  * the reference's AssembleBEMMatrix through the hooked library against the plain reference library, on the compact
    cases of unitTests/unit-test-BEMMatrix.cc:108-150 (PEC, dielectric, two-body with Mate reuse; real, small and imaginary
    frequency), <= 1e-10 relative Frobenius (north_star tolerance);
  * the reference's own scuff-scatter application (applications/scuff-scatter, built from the reference's sources against
    the hooked library) on the tests/Mie geometry: absorbed / scattered power, force and torque equal to the CPU run's to
    the precision the application prints.
The binaries are built where /root/reference exists and travel to the GPU box (oracle/_ref is git-ignored, not gpurun-ignored)."""
import os
import subprocess
import numpy as np
import pytest
from conftest import GEODIR, ROOT, rel_fro

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _need(*names):
    for n in names:
        if not os.path.exists(os.path.join(REFDIR, n)):
            pytest.skip("oracle/_ref/%s not built (make -C oracle cuda where /root/reference exists)" % n)


def _matrix(binary, geo, omega, imag, tmp, cuda):
    out = os.path.join(tmp, ("cuda" if cuda else "cpu") + ".bin")
    cmd = [os.path.join(REFDIR, binary), "matrix", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), out]
    if imag:
        cmd.append("--imag")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    env.pop("SCUFF_CUDA", None)
    if cuda:
        env.update(SCUFF_CUDA="1", SCUFF_CUDA_VERBOSE="1")
    r = subprocess.run(cmd, cwd=os.path.dirname(geo), capture_output=True, text=True, env=env, stdin=subprocess.DEVNULL, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    M = np.fromfile(out, dtype=np.float64 if imag else np.complex128)
    n = int(round(np.sqrt(M.size)))
    return M.reshape(n, n).T, r.stderr


CASES = [("PECSphere_255.scuffgeo", 0.01, False), ("PECSphere_255.scuffgeo", 1.0, False), ("PECSphere_255.scuffgeo", 0.1j, True),
         ("SiSphere_255.scuffgeo", 1.0, False), ("SiSphere_255.scuffgeo", 0.1j, True), ("SiSpheres_255.scuffgeo", 0.1, False),
         ("ZetaSphere_255.scuffgeo", 0.7, False)]


@pytest.mark.parametrize("geo,omega,imag", CASES)
def test_reference_library_through_the_binding(tmp_path, geo, omega, imag):
    _need("ref_driver", "ref_driver_cuda", "libscuff_ref_cuda.so")
    path = os.path.join(GEODIR, geo)
    ref, _ = _matrix("ref_driver", path, omega, imag, str(tmp_path), False)
    # the hooked library with SCUFF_CUDA unset IS the reference
    same, _ = _matrix("ref_driver_cuda", path, omega, imag, str(tmp_path), False)
    assert rel_fro(same, ref) <= 1e-12
    got, log = _matrix("ref_driver_cuda", path, omega, imag, str(tmp_path), True)
    assert "blocks assembled on the device" in log and " 0 BEM" not in log, log[-500:]
    err = rel_fro(got, ref)
    print(geo, omega, "reference API through libscuff_cuda: rel_fro", err, "|", log.strip().splitlines()[-1])
    assert err <= 1e-10


def test_scuff_scatter_through_the_binding(tmp_path):
    """the reference's scuff-scatter on tests/Mie/LossySphere_327 at two frequencies: EMTPFT output (absorbed and scattered
    power, force, torque) with SCUFF_CUDA=1 against the unmodified CPU run"""
    _need("scuff-scatter", "scuff-scatter_cuda")
    rows = {}
    for tag, binary, cuda in (("cpu", "scuff-scatter", False), ("cuda", "scuff-scatter_cuda", True)):
        d = tmp_path / tag; d.mkdir()
        for f in ("LossySphere_327.scuffgeo", "Sphere_327.msh"):
            os.symlink(os.path.join(GEODIR, f), str(d / f))
        env = dict(os.environ, OMP_NUM_THREADS="4")
        env.pop("SCUFF_CUDA", None)
        if cuda:
            env.update(SCUFF_CUDA="1", SCUFF_CUDA_VERBOSE="1")
        cmd = [os.path.join(REFDIR, binary), "--geometry", "LossySphere_327.scuffgeo", "--Omega", "0.4", "--Omega", "1.0", "--PWDirection", "0", "0", "1",
               "--PWPolarization", "0.70710678", "0.70710678i", "0", "--EMTPFTFile", "out.EMTPFT"]
        # (the reference's ProcessOptions also reads options from stdin when it is not a terminal: give it an empty one)
        r = subprocess.run(cmd, cwd=str(d), capture_output=True, text=True, env=env, stdin=subprocess.DEVNULL, timeout=300)
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        if cuda:
            assert "blocks assembled on the device" in r.stderr and " 0 BEM" not in r.stderr, r.stderr[-500:]
        rows[tag] = np.array([[float(x) for x in l.split()[2:]] for l in open(d / "out.EMTPFT") if not l.startswith("#")])
    assert rows["cpu"].shape == rows["cuda"].shape and rows["cpu"].shape[0] == 2
    scale = np.abs(rows["cpu"]).max(axis=1, keepdims=True)
    err = np.abs(rows["cuda"] - rows["cpu"]) / scale
    print("scuff-scatter EMTPFT (PAbs, PScat, F, T) cpu vs cuda, max deviation relative to the row's largest entry:", err.max())
    print(rows["cpu"][1], rows["cuda"][1])
    assert err.max() <= 2e-6          # the application prints 7 significant digits


def test_derivative_blocks_through_the_binding(tmp_path):
    """AssembleBEMMatrixBlock(GradM, dMdT) of the hooked reference library: SCUFF_CUDA=1 vs the same binary with it unset"""
    _need("ref_driver_cuda", "libscuff_ref_cuda.so")
    geo = os.path.join(GEODIR, "PECSpheres_327.scuffgeo")
    outs = []
    for cuda in (False, True):
        out = os.path.join(str(tmp_path), "g%d.bin" % cuda)
        cmd = [os.path.join(REFDIR, "ref_driver_cuda"), "gradblock", os.path.basename(geo), "0.0", "0.5", "0", "1", out, "--imag",
               "--torque", "0", "0", "1", "--torque", "1", "0", "0"]
        env = dict(os.environ, OMP_NUM_THREADS="8")
        env.pop("SCUFF_CUDA", None)
        if cuda:
            env.update(SCUFF_CUDA="1", SCUFF_CUDA_VERBOSE="1")
        r = subprocess.run(cmd, cwd=GEODIR, capture_output=True, text=True, check=True, env=env)
        if cuda:
            assert "blocks assembled on the device" in r.stderr
        outs.append(np.fromfile(out, dtype=np.float64).reshape(6, 327, 327))
    for q, lab in enumerate(["M", "dM/dx", "dM/dy", "dM/dz", "dM/dT0", "dM/dT1"]):
        err = np.linalg.norm(outs[1][q] - outs[0][q]) / np.linalg.norm(outs[0][q])
        print("hooked reference", lab, "rel err %.3e" % err)
        assert err <= 1e-10


def test_scuff_scatter_rhs_and_pft_through_the_binding(tmp_path):
    """SURVEY 8 row f2 through the reference's own application: with the cubature integrals selected (SCUFF_EMTPFTI_METHOD=0) the hooked
    scuff-scatter also assembles its right-hand side and evaluates the EMTPFT sums on the device (hooks in AssembleRHSVector and
    GetEMTPFTMatrix); output file against the unmodified CPU run with the same setting.  Two spheres: every surface-pair branch."""
    _need("scuff-scatter", "scuff-scatter_cuda")
    rows = {}
    for tag, binary, cuda in (("cpu", "scuff-scatter", False), ("cuda", "scuff-scatter_cuda", True)):
        d = tmp_path / tag; d.mkdir()
        for f in ("E10Spheres_327.scuffgeo", "Sphere_327.msh"):
            os.symlink(os.path.join(GEODIR, f), str(d / f))
        env = dict(os.environ, OMP_NUM_THREADS="8", SCUFF_EMTPFTI_METHOD="0")
        env.pop("SCUFF_CUDA", None)
        if cuda:
            env.update(SCUFF_CUDA="1", SCUFF_CUDA_VERBOSE="1")
        cmd = [os.path.join(REFDIR, binary), "--geometry", "E10Spheres_327.scuffgeo", "--Omega", "0.8", "--PWDirection", "0", "0", "1",
               "--PWPolarization", "1", "0", "0", "--EMTPFTFile", "out.EMTPFT"]
        r = subprocess.run(cmd, cwd=str(d), capture_output=True, text=True, env=env, stdin=subprocess.DEVNULL, timeout=600)
        assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
        if cuda:
            assert "1 right-hand sides, 1 EMTPFT evaluations on the device" in r.stderr, r.stderr[-500:]
        rows[tag] = np.array([[float(x) for x in l.split()[2:]] for l in open(d / "out.EMTPFT") if not l.startswith("#")])
    assert rows["cpu"].shape == rows["cuda"].shape and rows["cpu"].shape[0] == 2      # one row per sphere
    scale = np.abs(rows["cpu"]).max(axis=1, keepdims=True)
    err = np.abs(rows["cuda"] - rows["cpu"]) / scale
    print("scuff-scatter (RHS + EMTPFT on the device) cpu vs cuda:", err.max()); print(rows["cpu"], rows["cuda"])
    assert err.max() <= 2e-6          # the application prints 7 significant digits
