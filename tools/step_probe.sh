#!/bin/bash
# whole-step experiments (far + near + gather) on one GPU box:  gpurun --timeout 900 -- bash tools/step_probe.sh [tag]
TAG=${1:-probe}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_step_probe.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py --nu 26 --reps 4 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  total %.3f  far %.3f  near %.3f  gather %.3f' % (d['ms_total'], d['ms_far'], d['ms_near'], d['ms_gather']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
run SCUFF_CUDA_FAR_GATHER_RED=0
run SCUFF_CUDA_FAR_GATHER_RED=1
run SCUFF_CUDA_FAR_GATHER_RED=1 SCUFF_CUDA_TD_GROUP=16
run SCUFF_CUDA_FAR_GATHER_RED=1 SCUFF_CUDA_TD_GROUP=32
run SCUFF_CUDA_FAR_GATHER_RED=0 SCUFF_CUDA_TD_GROUP=32
run SCUFF_CUDA_FAR_TWOPASS=0
echo "== dielectric nu=18" >> $OUT
rund() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py --nu 18 --material CONST_EPS_10 --reps 4 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  total %.3f  far %.3f  near %.3f  gather %.3f' % (d['ms_total'], d['ms_far'], d['ms_near'], d['ms_gather']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
rund SCUFF_CUDA_FAR_GATHER_RED=0
rund SCUFF_CUDA_FAR_GATHER_RED=1
rund SCUFF_CUDA_FAR_TWOPASS=0
cat $OUT
