#!/bin/bash
# Gradient path session: parity tests, ContractGrads benches (CPU restatement beside it for h2o/bo3h3), ncu launch list.
TAG=${TAG:-g}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
for w in ${GRAD_WORKLOADS:-h2o bo3h3}; do
  timeout 900 python bench.py --path grad --workload $w --steps 5 --warmup 2 > gpurun_out/${TAG}_grad_$w.json 2> gpurun_out/${TAG}_grad_$w.err
  echo "grad $w rc=$?"; head -c 700 gpurun_out/${TAG}_grad_$w.json; echo
done
for w in ${GRAD_BIG:-}; do
  timeout 900 python bench.py --path grad --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_grad_$w.json 2> gpurun_out/${TAG}_grad_$w.err
  echo "grad $w rc=$?"; head -c 700 gpurun_out/${TAG}_grad_$w.json; echo
done
if [ -n "$NCU_LIST" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_grad_launches_$NCU_LIST.csv \
    python bench.py --path grad --workload $NCU_LIST --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_grad_ncu_list.log 2>&1
  echo "ncu list rc=$?"
fi
for w in ${WORKLOADS:-}; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; head -c 300 gpurun_out/${TAG}_bench_$w.json; echo
done
