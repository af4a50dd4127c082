#!/bin/bash
# A/B of gradient kernel builds: VARIANTS="base v1" GRAD_WORKLOADS="bo3h3 fe4s4 c18"
TAG=${TAG:-gab}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
for v in ${VARIANTS:-base}; do
  lib=$PWD/chinium_b200/libchinium_fock_$v.so; [ "$v" = base ] && lib=$PWD/chinium_b200/libchinium_fock.so
  for w in ${GRAD_WORKLOADS:-bo3h3}; do
    CHINIUM_FOCK_LIB=$lib timeout 900 python bench.py --path grad --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_${v}_grad_$w.json 2> gpurun_out/${TAG}_${v}_grad_$w.err
    echo "grad $v $w rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_${v}_grad_$w.json')); print('  ms %.2f frac %.4f' % (d['ms_per_step'], d['roofline']['frac']))"
  done
done
