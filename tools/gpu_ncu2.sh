#!/bin/bash
# ncu --set full captures: NAME WORKLOAD REGEX COUNT (repeatable via args in groups of 4)
mkdir -p gpurun_out
while [ $# -ge 4 ]; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$3" -c "$4" \
    -f -o gpurun_out/$1 python bench.py --workload $2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/$1.log 2>&1
  echo "ncu $1 rc=$?"; tail -2 gpurun_out/$1.log; ls -la gpurun_out/$1.ncu-rep
  shift 4
done
