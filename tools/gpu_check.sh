#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (+ reference arm), ncu launch list and a full capture of the far kernel.
#   gpurun --timeout 2400 -- bash tools/gpu_check.sh [tag]
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-lu --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:far_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_far \
   python bench.py --steps 1 --warmup 3 --no-lu --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_ncu_far.log 2>&1; echo "ncu full rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log | tail -3; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
