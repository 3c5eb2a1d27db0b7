"""Golden values of the reference's Ewald-summed periodic Green's function GBarVDEwald (GBarVDEwald.cc:931-996), written by the
UNMODIFIED reference through oracle/_ref/ref_driver gbar.  Seed data for the periodic (LATTICE) outer-cell step.
Run in the build container:   python tests/golden/make_gbar_golden.py"""
import json, os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
DRV = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "ref_driver")
CASES = [dict(k=[0.0, 2.0], kb=[0.3, -0.2], ldim=2, lbv=[[1, 0], [0, 1]], excl=0),
         dict(k=[0.0, 2.0], kb=[0.3, -0.2], ldim=2, lbv=[[1, 0], [0, 1]], excl=1),
         dict(k=[0.5, 1.5], kb=[1.0, 0.4], ldim=2, lbv=[[1, 0], [0, 0.8]], excl=1),
         dict(k=[0.2, 2.5], kb=[0.7, 0.0], ldim=1, lbv=[[0.75, 0], [0, 0]], excl=0),
         dict(k=[1.0, 0.0], kb=[0.3, -0.2], ldim=2, lbv=[[1, 0], [0, 1]], excl=1),          # real k: no direct-sum check possible
         dict(k=[0.5, 1.5], kb=[1.0, 0.4], ldim=2, lbv=[[1, 0], [0.3, 0.8]], excl=0, reference_not_converged=True)]   # oblique lattice
out = []
for i, c in enumerate(CASES):
    fn = "/tmp/gbar_golden.bin"
    subprocess.run([DRV, "gbar", repr(float(c["k"][0])), repr(float(c["k"][1])), repr(float(c["kb"][0])), repr(float(c["kb"][1])), str(c["ldim"]),
                    repr(float(c["lbv"][0][0])), repr(float(c["lbv"][0][1])), repr(float(c["lbv"][1][0])), repr(float(c["lbv"][1][1])),
                    str(c["excl"]), "6", str(100 + i), fn], check=True)
    a = np.fromfile(fn).reshape(-1, 19)
    c = dict(c); c["R"] = a[:, :3].tolist(); c["GBarVD"] = a[:, 3:].tolist(); out.append(c)
json.dump(out, open(os.path.join(HERE, "gbar_golden.json"), "w"))
