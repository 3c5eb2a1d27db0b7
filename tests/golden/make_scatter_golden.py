"""Golden vectors of the scattering-run rows from the UNMODIFIED reference (oracle/_ref/ref_driver scatter).
Run in the build container:   python tests/golden/make_scatter_golden.py   ->  tests/golden/scatter_golden.npz

Per case (geometry, omega; plane wave E0 = x, propagating along +z): the reference's right-hand side, solution vector, PFT
matrix by the energy/momentum-transfer method with the cubature integrals (EMTPFTIMethod 0), PFT matrix by the displaced-surface-
integral method with its defaults (302-point Lebedev rule, automatic radius), and total fields at 12 seeded evaluation points."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest

CASES = [("PECSphere_501", 1.0), ("SiO2Sphere_501", 1.0), ("E10Spheres_327", 0.8)]
PW = ((1.0, 0.0, 0.0), (0.0, 0.0, 1.0))


def points(name):
    rng = np.random.default_rng(7)
    d = rng.standard_normal((12, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    c = np.array([0.0, 0.0, 1.5]) if "Spheres" in name else np.zeros(3)
    r = rng.uniform(3.5, 6.0, 12) if "Spheres" in name else rng.uniform(1.6, 5.0, 12)
    return c + d * r[:, None]


def main():
    out = {}
    for name, omega in CASES:
        geo = os.path.join(HERE, "geometries", name + ".scuffgeo")
        X = points(name)
        RHS, KN, PFT, info = conftest.ref_scatter(geo, omega, PW, points=X, dsi=(0.0, 302), tmpdir="/tmp/scuff_golden_scatter")
        key = "%s@%g" % (name, omega)
        out[key + ":rhs"] = RHS; out[key + ":kn"] = KN; out[key + ":emtpft0"] = PFT
        out[key + ":dsipft"] = info["dsipft"]; out[key + ":x"] = X; out[key + ":fields"] = info["fields"]
        print(key, PFT[0, :2], info["dsipft"][0, :2])
    np.savez_compressed(os.path.join(HERE, "scatter_golden.npz"), **out)


if __name__ == "__main__":
    main()
