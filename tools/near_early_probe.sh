#!/bin/bash
# step time with the near/singular kernels started after the first far pass (default) or together with it
#   gpurun --timeout 600 -- bash tools/near_early_probe.sh [tag]
TAG=${1:-probe}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_near_early_probe.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py $ARGS --reps 5 2>&1 | tail -3 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  total %.3f  far %.3f  near %.3f  gather %.3f  wall %.3f' % (d['ms_total'], d['ms_far'], d['ms_near'], d['ms_gather'], 1e3*d['wall_s']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
ARGS="--nu 26"; run X=1; run SCUFF_CUDA_NEAR_EARLY=1
ARGS="--nu 18 --material CONST_EPS_10"; run X=1; run SCUFF_CUDA_NEAR_EARLY=1
ARGS="--nu 26 --omega 1j --imag"; run X=1; run SCUFF_CUDA_NEAR_EARLY=1
cat $OUT
