"""Derivative blocks GradM[3] / dMdT[NumTorqueAxes] of AssembleBEMMatrixBlock (scuffcuda_assemble_block_gradients, grad_kernel.cu)
against the UNMODIFIED reference's AssembleBEMMatrixBlock with the same arguments (ref_driver gradblock): the blocks scuff-cas3D
asks for on the U blocks of different objects (CasimirIntegrand.cc:443-447) — PEC and dielectric spheres, real, imaginary and
complex frequency, real matrices at imaginary frequency, transformed geometries (the .trans file of the Casimir test), torque
axes; far pairs (6x6 rule) and, for the touching-distance transformation, near pairs (78x78 rule)."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_gradblock

pytestmark = pytest.mark.gpu

AXES = [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (0.3, -0.5, 0.8)]


def device_blocks(ps, geo, omega, real, axes, trans=None):
    G = ps.RWGGeometry(geo)
    if trans is not None:
        G.transform(trans)
    sa, sb = G.surface(0), G.surface(1)
    mk = lambda: ps.HMatrix.new(sa["NBF"], sb["NBF"], real)
    M = mk(); GradM = [mk(), mk(), mk()]; dMdT = [mk() for _ in axes]
    G.AssembleBEMMatrixBlockGrad(0, 1, omega, M, GradM=GradM, TorqueAxes=list(axes), dMdT=dMdT)
    out = [M.array.copy()] + [g.array.copy() for g in GradM] + [t.array.copy() for t in dMdT]
    if trans is not None:
        G.untransform()
    return out


@pytest.mark.parametrize("name,omega,real,ntrans", [("PECSpheres_327", 0.5j, True, None), ("PECSpheres_327", 0.5j, True, 3),
                                                    ("PECSpheres_327", 0.7, False, None), ("E10Spheres_327", 0.3, False, None),
                                                    ("SiSpheres_255", 0.25j, False, None), ("PECSpheres_255", 0.4 + 0.2j, False, None)])
def test_gradient_and_torque_blocks(ps, name, omega, real, ntrans):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    trans, transline = None, None
    if ntrans is not None:
        tf = os.path.join(GEODIR, "Spheres.trans")
        trans = ("Spheres.trans", ntrans)
        transline = [l for l in open(tf) if l.strip() and not l.strip().startswith("#")][ntrans]
    ref, info = ref_gradblock(geo, omega, 0, 1, imag=real, torque_axes=AXES, trans=trans, env={"OMP_NUM_THREADS": "8"})
    dev = device_blocks(ps, geo, omega, real, AXES, transline)
    labels = ["M", "dM/dx", "dM/dy", "dM/dz"] + ["dM/dT%d" % i for i in range(len(AXES))]
    for lab, A, B in zip(labels, dev, ref):
        assert A.dtype == B.dtype and A.shape == B.shape
        err = np.linalg.norm(A - B) / np.linalg.norm(B)
        print(name, omega, ntrans, lab, "rel err %.3e  |ref| %.3e" % (err, np.linalg.norm(B)))
        assert err <= 1e-10


def test_gradients_close_objects(ps, tmp_path):
    """two spheres 0.05 apart: panel pairs with rRel <= 4 between the objects take the 78x78 rule for the derivatives"""
    import shutil
    shutil.copy(os.path.join(GEODIR, "SSphere_255.msh"), tmp_path)
    geo = str(tmp_path / "Close.scuffgeo")
    open(geo, "w").write("OBJECT A\n\tMESHFILE SSphere_255.msh\nENDOBJECT\nOBJECT B\n\tMESHFILE SSphere_255.msh\n\tDISPLACED 0 0 2.05\nENDOBJECT\n")
    for omega, real in ((0.8j, True), (1.3, False)):
        ref, _ = ref_gradblock(geo, omega, 0, 1, imag=real, torque_axes=AXES[:1], env={"OMP_NUM_THREADS": "8"})
        dev = device_blocks(ps, geo, omega, real, AXES[:1])
        for lab, A, B in zip(["M", "dM/dx", "dM/dy", "dM/dz", "dM/dT0"], dev, ref):
            err = np.linalg.norm(A - B) / np.linalg.norm(B)
            print("close spheres", omega, lab, "rel err %.3e" % err)
            assert err <= 1e-10


def test_gradients_refused_where_not_available(ps):
    G = ps.RWGGeometry(os.path.join(GEODIR, "PECSpheres_255.scuffgeo"))
    n = G.surface(0)["NBF"]
    M = ps.HMatrix.new(n, n, False)
    with pytest.raises(ps.ScuffError, match="two different"):
        G.AssembleBEMMatrixBlockGrad(0, 0, 1.0, M, GradM=[ps.HMatrix.new(n, n, False), None, None])
