"""Device FIPPI builder (adaptive-cubature replica) parity: the table is built on the GPU from scratch
(no preloaded cache) and (a) its records are compared with the records the UNMODIFIED reference wrote to
its cache file, (b) the full matrix assembled from it is compared with the reference (<= 1e-10)."""
import os
import numpy as np
import pytest
from conftest import GEODIR, ref_matrix

pytestmark = pytest.mark.gpu
REC = np.dtype([("key", "<f4", 15), ("pad", "<f4"), ("data", "<f8", 36)])


def read_cache(fn):
    raw = open(fn, "rb").read()
    assert raw[:11] == b"FIPPICACHE\0"
    r = np.frombuffer(raw[11:], dtype=REC)
    return {x["key"].tobytes(): x["data"] for x in r}


@pytest.mark.parametrize("name,omega", [("PECSphere_501", 1.0), ("SiSpheres_255", 0.1), ("PECSphere_R0P75_414", 0.3 + 0.1j)])
def test_fippi_builder(ps, name, omega, tmp_path):
    geo = os.path.join(GEODIR, name + ".scuffgeo")
    refcache, mycache = str(tmp_path / "ref.cache"), str(tmp_path / "mine.cache")
    Mref, _ = ref_matrix(geo, omega, cache_out=refcache)
    ps.ClearCache()
    G = ps.RWGGeometry(geo)
    M = G.AssembleBEMMatrix(omega).array
    t = ps.last_timing()
    ps.StoreCache(mycache)
    ref, mine = read_cache(refcache), read_cache(mycache)
    common = set(ref) & set(mine)
    assert len(common) == len(ref), "every record the reference needed must have been built"
    worst, nexact = 0.0, 0
    for k in common:
        a, b = mine[k][:33], ref[k][:33]
        worst = max(worst, np.abs(a - b).max() / np.abs(b).max())
        nexact += int(np.array_equal(mine[k], ref[k]))
    err = np.linalg.norm(M - Mref) / np.linalg.norm(Mref)
    print(name, omega, "records ref/mine", len(ref), len(mine), "bit-exact", nexact, "worst rel record diff", worst,
          "matrix rel err", err, "timing(ms) total/far/near/fippi", t[:4])
    assert worst <= 1e-9
    assert err <= 1e-10
