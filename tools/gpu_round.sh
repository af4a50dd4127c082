#!/bin/bash
# One scripted GPU session: parity tests, benches, per-class timing, ncu launch list. Outputs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep 'Model name' >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in h2o bo3h3 fe4s4 c18; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --per-class > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; head -c 600 gpurun_out/bench_$w.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bo3h3.csv python bench.py --workload bo3h3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bo3h3.log 2>&1
echo "ncu rc=$?"
