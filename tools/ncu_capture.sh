#!/bin/bash
# usage: tools/ncu_capture.sh <name> <workload> <kernel-regex> <count>
# One `ncu --set full` capture of selected class-pair kernels of one J/K build (1 GPU). Report -> gpurun_out/<name>.ncu-rep
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$3" -c "$4" \
  -f -o gpurun_out/$1 python bench.py --workload $2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/$1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/$1.log
