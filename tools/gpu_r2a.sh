#!/bin/bash
# Round-2 session A: new parity tests (BASELINE configs), the round-1 suite, benches with the executed-work roofline.
TAG=${TAG:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep 'Model name' >> gpurun_out/${TAG}_gpu.txt; free -g >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_baseline.py -m gpu -q -s --durations=0 > gpurun_out/${TAG}_pytest_baseline.log 2>&1; echo "pytest baseline rc=$?"
tail -25 gpurun_out/${TAG}_pytest_baseline.log
timeout 1200 python -m pytest tests/test_gpu.py -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -15 gpurun_out/${TAG}_pytest_gpu.log
for w in ${WORKLOADS:-c18 fe4s4 h2o64}; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --per-class > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; head -c 300 gpurun_out/${TAG}_bench_$w.json; echo; tail -3 gpurun_out/${TAG}_bench_$w.err
done
