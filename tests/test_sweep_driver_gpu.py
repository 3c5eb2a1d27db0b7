"""scuff::SweepDriver (C++: one host thread + device context per worker, FIPPI table built once by the workers together) on the
reference's Mie sweep (tests/Mie/OmegaFile on SiO2Sphere_501): 2 workers against 1 worker.  The workers share one GPU here (the
test box has one); `tests/sweep_driver_test <geo> <OmegaFile> 2 0 1` runs them on two GPUs."""
import json
import os
import subprocess
import pytest
from conftest import GEODIR, ROOT

pytestmark = pytest.mark.gpu


def test_sweep_driver_two_workers(tmp_path):
    exe = str(tmp_path / "sweep_driver_test")
    lib = os.path.join(ROOT, "scuff-em_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "sweep_driver_test.cc"), "-o", exe, "-L", lib, "-lscuff_b200", "-lscuff_cuda",
                    "-Wl,-rpath," + lib, "-lpthread"], check=True)
    import torch
    devs = ["0", "1"] if torch.cuda.device_count() > 1 else ["0", "0"]
    r = subprocess.run([exe, "SiO2Sphere_501.scuffgeo", "OmegaFile", "2"] + devs, cwd=GEODIR, capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["points"] == 15 and out["workers"] == 2 and out["points_worker0"] == 8
    assert out["worst_rel_diff"] <= 1e-9
