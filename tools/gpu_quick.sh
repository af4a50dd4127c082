#!/bin/bash
# quick GPU check: parity tests + c18/fe4s4 bench with per-class table
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c18 fe4s4}; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w rc=$?"; head -c 300 gpurun_out/bench_$w.json; echo
done
