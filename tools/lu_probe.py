import sys, os, time, importlib.util
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "scuff-em_b200"))
import meshgen
spec = importlib.util.spec_from_file_location("pyscuff", os.path.join(ROOT, "scuff-em_b200", "pyscuff.py"))
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)
geo = meshgen.write_sphere_geo("/tmp/scuff_b200_mesh", 26, None)
G = ps.RWGGeometry(geo); M = G.AllocateBEMMatrix(); N = G.TotalBFs
G.AssembleBEMMatrix(1.0, M=M, to_host=False)
X = ps.HMatrix.new(N, 1, False)
def lu(tag):
    t0=time.perf_counter(); M.lu_factorize(); t1=time.perf_counter(); M.lu_solve(X); t2=time.perf_counter()
    print(tag, "factor %.3f solve %.3f" % (t1-t0, t2-t1), flush=True)
lu("after device assembly (first)")
G.AssembleBEMMatrix(1.0, M=M, to_host=False); lu("after device assembly (second)")
mode = os.environ.get("MODE","upper")
os.environ["SCUFF_CUDA_PIPELINE_MODE"] = mode
t0=time.perf_counter(); G.AssembleBEMMatrix(1.0, M=M, to_host=True); print("e2e", mode, time.perf_counter()-t0)
lu("after host copy-back (%s)" % mode)
G.AssembleBEMMatrix(1.0, M=M, to_host=False); lu("after device assembly again")
