#!/bin/bash
# ncu --set full capture of the PFT pair kernel (rhs_pft.cu) at N = 20 280 (PEC icosphere, same-surface variant) and at the
# dielectric nu=18 icosphere (magnetic currents), text summaries only.   gpurun --timeout 900 -- bash tools/profile_pft.sh r02
TAG=${1:-r02}
mkdir -p gpurun_out
for cfg in "pec_N20280 --nu 26" "diel_N19440 --nu 18 --material CONST_EPS_10"; do
  set -- $cfg; name=$1; shift
  python tools/run_assemble.py "$@" --pft --reps 3 > gpurun_out/${TAG}_pft_${name}.timing.txt 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pft_pair_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_pft_pair_${name} \
     python tools/run_assemble.py "$@" --pft --reps 2 > gpurun_out/${TAG}_ncu_pft_${name}.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py gpurun_out/${TAG}_pft_pair_${name}.ncu-rep > gpurun_out/${TAG}_pft_pair_${name}.summary.txt 2>&1
  python tools/ncu_srcprof.py gpurun_out/${TAG}_pft_pair_${name}.ncu-rep 30 200 > gpurun_out/${TAG}_pft_pair_${name}.srcprof.txt 2>&1
  rm -f gpurun_out/${TAG}_pft_pair_${name}.ncu-rep
done
cat gpurun_out/${TAG}_pft_*.timing.txt
