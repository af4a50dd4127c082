#!/bin/bash
# One scripted GPU session: parity tests, c18/fe4s4 benches with per-class table, ncu launch list, ncu --set full
# captures of representative kernels.  Outputs in gpurun_out/<TAG>_*.
TAG=${TAG:-s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep 'Model name' >> gpurun_out/${TAG}_gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
for w in ${WORKLOADS:-c18 fe4s4}; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --per-class $BENCH_FLAGS > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; head -c 400 gpurun_out/${TAG}_bench_$w.json; echo
done
if [ -n "$NCU_LIST" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$NCU_LIST.csv \
    python bench.py --workload $NCU_LIST --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
  echo "ncu list rc=$?"
fi
if [ -n "$NCU_FULL" ]; then   # NCU_FULL="<workload>:<regex>:<count>"
  IFS=: read W RX CNT <<< "$NCU_FULL"
  timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$RX" -c "$CNT" \
    -f -o gpurun_out/${TAG}_full python bench.py --workload $W --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_ncu_full.log
fi
