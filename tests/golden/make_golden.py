"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference).  Run in the build container (needs oracle/_ref):   python tests/golden/make_golden.py

  geom/<name>.geom.txt        RWGGeometry as the reference parsed it (vertices, panels, edge order, BF offsets)
  ppi_golden.json             GetPanelPanelInteractions H[0],H[1] + algorithm id for a seeded set of panel pairs
                              of PECSphere_R0P75_414 (the geometry of unitTests/unit-test-PPIs.cc) at k=0.1, 5, 5+20i
  matrix_samples.npz          per (geometry, omega): ||M||_F and 4000 seeded entries (row, col, value) of the
                              reference's AssembleBEMMatrix output
  eps_golden.json             MatProp values the reference printed
"""
import json, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GEODIR = os.path.join(HERE, "geometries")
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
NAMES = ["PECSphere_501", "PECSphere_255", "PECSphere_R0P75_414", "PECSpheres_255", "PECSpheres_327", "SiO2Sphere_501",
         "GoldSphere_501", "SiSphere_255", "SiSpheres_255", "LossySphere_327", "E10Spheres_327", "PECPlate_40", "SiSlab_40",
         "SphereSlabArray", "OpenPlate_40"]
CASES = [("PECSphere_501", 1.0, False), ("PECSphere_501", 0.01, False), ("PECSphere_255", 0.1j, False),
         ("PECSphere_255", 0.1j, True), ("PECSphere_255", 1.0, False), ("SiO2Sphere_501", 1.0, False),
         ("SiSphere_255", 0.01, False), ("PECSpheres_255", 0.1, False), ("SiSpheres_255", 0.1, False),
         ("LossySphere_327", 1.0, False), ("PECSpheres_327", 1.0j, True), ("E10Spheres_327", 0.5j, False),
         ("GoldSphere_501", 0.3, False), ("PECSphere_501", 0.7 + 0.3j, False)]


def run(args):
    return subprocess.run([DRV] + args, cwd=GEODIR, capture_output=True, text=True, check=True).stdout


def main():
    os.makedirs(os.path.join(HERE, "geom"), exist_ok=True)
    for n in NAMES:
        run(["geom", n + ".scuffgeo", os.path.join(HERE, "geom", n + ".geom.txt")])
    # ---- RWG overlap integrals (surface-impedance term) ----
    for n in ("ZetaSphere_255", "OpenPlate_40"):
        run(["overlaps", n + ".scuffgeo", "0", os.path.join(HERE, "overlaps_%s.txt" % n)])
    # ---- PPIs ----
    rng = np.random.default_rng(20261017)
    pairs = [(0, 0, 0, 0, 223, 0), (0, 0, 0, 0, 65, 0), (0, 0, 0, 0, 66, 0), (0, 0, 0, 0, 0, 0)]   # unit-test-PPIs.cc cases
    for _ in range(60):
        pairs.append((0, int(rng.integers(276)), int(rng.integers(3)), 0, int(rng.integers(276)), int(rng.integers(3))))
    open("/tmp/ppi_pairs.txt", "w").write("".join("%d %d %d %d %d %d\n" % p for p in pairs))
    out = {"geometry": "PECSphere_R0P75_414", "pairs": pairs, "k": [], "H": []}
    for k in (0.1, 5.0, 5.0 + 20.0j):
        run(["ppis", "PECSphere_R0P75_414.scuffgeo", repr(k.real), repr(complex(k).imag), "/tmp/ppi_pairs.txt", "/tmp/ppi_out.bin"])
        out["k"].append([complex(k).real, complex(k).imag])
        out["H"].append(np.fromfile("/tmp/ppi_out.bin").reshape(-1, 5).tolist())
    json.dump(out, open(os.path.join(HERE, "ppi_golden.json"), "w"))
    # ---- matrix samples ----
    S = {}
    epsfile = os.path.join(HERE, "eps_golden.json")
    EPS = json.load(open(epsfile)) if os.path.exists(epsfile) else {}
    for (name, omega, imag) in CASES:
        args = ["matrix", name + ".scuffgeo", repr(complex(omega).real), repr(complex(omega).imag), "/tmp/m.bin"] + (["--imag"] if imag else [])
        info = json.loads(run(args).strip().splitlines()[-1])
        M = np.fromfile("/tmp/m.bin", dtype=np.float64 if imag else np.complex128).reshape(info["NC"], info["NR"]).T
        key = "%s@%r@%d" % (name, complex(omega), int(imag))
        if len(info["eps"]) > 1 and complex(omega).imag == 0:
            EPS["%s@%g" % (name, complex(omega).real)] = info["eps"][1]
        idx = rng.integers(0, M.shape[0], size=(4000, 2))
        S[key + "/idx"] = idx.astype(np.int32)
        S[key + "/val"] = M[idx[:, 0], idx[:, 1]].astype(np.complex128)
        S[key + "/fro"] = np.array([np.linalg.norm(M), M.shape[0]])
        print(key, "fro", np.linalg.norm(M))
    json.dump(EPS, open(epsfile, "w"), indent=0)
    np.savez_compressed(os.path.join(HERE, "matrix_samples.npz"), **S)


if __name__ == "__main__":
    main()
