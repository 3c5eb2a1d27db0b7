#!/bin/bash
# far-kernel experiments on one GPU box: kernel time alone (far-only assemblies) for several launch configurations
#   gpurun --timeout 900 -- bash tools/far_probe.sh [tag]
TAG=${1:-probe}
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_far_probe.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py --nu 26 --far-only --reps 4 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  ms_far %.3f  gather %.3f  total %.3f' % (d['ms_far'], d['ms_gather'], d['ms_total']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
run SCUFF_CUDA_FAR_TWOPASS=0 SCUFF_CUDA_FAR_STAGED=0 SCUFF_CUDA_FAR_SMALL=0
run SCUFF_CUDA_FAR_TWOPASS=0 SCUFF_CUDA_FAR_STAGED=1 SCUFF_CUDA_FAR_SMALL=0
run SCUFF_CUDA_FAR_TWOPASS=0 SCUFF_CUDA_FAR_STAGED=1 SCUFF_CUDA_FAR_SMALL=-1
run SCUFF_CUDA_FAR_TWOPASS=1 SCUFF_CUDA_FAR_SMALL=0
run SCUFF_CUDA_FAR_TWOPASS=1 SCUFF_CUDA_FAR_SMALL=1
run SCUFF_CUDA_FAR_TWOPASS=1 SCUFF_CUDA_FAR_SMALL=-1
echo "== imaginary frequency, real matrix" >> $OUT
runi() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py --nu 26 --far-only --reps 4 --omega 1j --imag 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  ms_far %.3f  gather %.3f  total %.3f' % (d['ms_far'], d['ms_gather'], d['ms_total']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
runi SCUFF_CUDA_FAR_TWOPASS=0 SCUFF_CUDA_FAR_STAGED=1
runi SCUFF_CUDA_FAR_TWOPASS=1
echo "== dielectric nu=18" >> $OUT
rund() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py --nu 18 --material CONST_EPS_10 --far-only --reps 4 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  ms_far %.3f  gather %.3f  total %.3f' % (d['ms_far'], d['ms_gather'], d['ms_total']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
rund SCUFF_CUDA_FAR_SMALL=0 SCUFF_CUDA_FAR_MINB_PSI=2
rund SCUFF_CUDA_FAR_SMALL=-1 SCUFF_CUDA_FAR_MINB_PSI=2
cat $OUT
