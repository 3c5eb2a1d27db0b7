#!/bin/bash
# timing experiments on the staged far kernel (SCUFF_CUDA_FAR_DBG: 1 no bulk ops, 2 plain bulk stores, 4 lanes 0..2 issue one op each)
TAG=${1:-probe}
OUT=gpurun_out/${TAG}_far_dbg.txt
: > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python tools/run_assemble.py --nu 26 --far-only --reps 3 2>>$OUT | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('  ms_far %.3f' % (d['ms_far']))
    except Exception: print(l.strip()[:200])" >> $OUT; }
for small in 0 -1; do for dbg in 16 80 82; do run SCUFF_CUDA_FAR_SMALL=$small SCUFF_CUDA_FAR_DBG=$dbg; done; done
cat $OUT
