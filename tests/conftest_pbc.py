"""Helpers around the reference's periodic Green's function (oracle/_ref/ref_driver gbarpts / gba).  Test infrastructure."""
import json
import os
import subprocess
import numpy as np
import pytest
from conftest import REF_DRIVER, have_ref


def ref_gbar_points(k, kb, lbv, excl, R, tmp):
    """GBarVDEwald of the UNMODIFIED reference at points R[n,3] (2-D lattice): complex [n,8]"""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    pin, pout = os.path.join(tmp, "p.bin"), os.path.join(tmp, "o.bin")
    np.ascontiguousarray(R, dtype=np.float64).tofile(pin)
    subprocess.run([REF_DRIVER, "gbarpts", repr(float(np.real(k))), repr(float(np.imag(k))), repr(float(kb[0])), repr(float(kb[1])),
                    repr(float(lbv[0][0])), repr(float(lbv[0][1])), repr(float(lbv[1][0])), repr(float(lbv[1][1])), str(int(excl)), pin, pout],
                   check=True)
    o = np.fromfile(pout).reshape(-1, 16)
    return o[:, 0::2] + 1j * o[:, 1::2]


def ref_gba(geo, omega, kb, nr, nsa, nsb, R, tmp):
    """CreateRegionGBA of the reference: (info dict or None, CTable, GetGBar values [n,4])"""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    pin, pout = os.path.join(tmp, "p.bin"), os.path.join(tmp, "g.bin")
    np.ascontiguousarray(R, dtype=np.float64).tofile(pin)
    r = subprocess.run([REF_DRIVER, "gba", os.path.basename(geo), repr(float(np.real(omega))), repr(float(np.imag(omega))), repr(float(kb[0])),
                        repr(float(kb[1])), str(nr), str(nsa), str(nsb), pin, pout], cwd=os.path.dirname(os.path.abspath(geo)),
                       capture_output=True, text=True, check=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
    J = json.loads(r.stdout.strip().splitlines()[-1])
    if J["null"]:
        return None, None, None
    ct = np.fromfile(pout + ".ctable") if os.path.exists(pout + ".ctable") and "N" in J else np.zeros(0)
    g = np.fromfile(pout).reshape(-1, 8)
    return J, ct, g[:, 0::2] + 1j * g[:, 1::2]
