#!/bin/bash
# ncu --set full captures of every hot kernel of the assembly path (one launch each) + the launch list of a bench step.
#   gpurun --timeout 2400 -- bash tools/profile_all.sh r02
TAG=${1:-r02}
mkdir -p gpurun_out
cap() { # name kernel-regex skip extra-env... -- command
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 env "$@" ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/${TAG}_$name \
     python tools/run_assemble.py $RUNARGS > gpurun_out/${TAG}_ncu_$name.log 2>&1; echo "$name rc=$?"
  # the reports are ~22 MB each and only 64 MB travel back: keep text summaries (metrics + source-level stall profile)
  python tools/ncu_summary.py gpurun_out/${TAG}_$name.ncu-rep > gpurun_out/${TAG}_$name.summary.txt 2>&1
  python tools/ncu_srcprof.py gpurun_out/${TAG}_$name.ncu-rep 30 200 > gpurun_out/${TAG}_$name.srcprof.txt 2>&1
  case $name in far_pass1_N20280|far_gather_N20280) ;; *) rm -f gpurun_out/${TAG}_$name.ncu-rep ;; esac; }
RUNARGS="--nu 26 --reps 2 --far-only"; cap far_pass1_N20280 far_pass1_kernel 1 X=1
RUNARGS="--nu 26 --reps 2 --far-only"; cap far_gather_N20280 far_gather_kernel 1 X=1
RUNARGS="--nu 26 --reps 2 --far-only"; cap far_pass1_general_N20280 far_pass1_kernel 1 SCUFF_CUDA_FAR_SMALL=0
RUNARGS="--nu 26 --reps 2 --far-only --omega 1j --imag"; cap far_pass1_imag_N20280 far_pass1_kernel 1 X=1
RUNARGS="--nu 18 --material CONST_EPS_10 --reps 2 --far-only"; cap far_pass1_psi_N19440 far_pass1_psi_kernel 2 X=1
RUNARGS="--nu 18 --material CONST_EPS_10 --reps 2 --far-only"; cap far_gather_psi_N19440 far_gather_psi_kernel 2 X=1
RUNARGS="--nu 26 --reps 2"; cap td_ct_pair_N20280 td_ct_pair_kernel 1 X=1
RUNARGS="--nu 26 --reps 2"; cap near_desing_N20280 near_desing_kernel 1 X=1
RUNARGS="--nu 26 --reps 2 --far-only"; cap far_staged_N20280 far_staged_kernel 1 SCUFF_CUDA_FAR_TWOPASS=0
RUNARGS="--nu 10 --reps 1 --omega 30"; cap ho_N3000 ho_kernel 0 X=1
RUNARGS="--nu 10 --reps 1"; cap fippi_coop_N3000 fippi_coop_kernel 0 X=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-lu --no-cpu-baseline --no-configs --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/${TAG}_*
