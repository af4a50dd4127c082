#!/bin/bash
# A/B: bench --per-class for several builds of the library.  VARIANTS="'' _v3 _v4"  WORKLOADS="c18 fe4s4"
TAG=${TAG:-ab}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
for v in ${VARIANTS:-base}; do
  lib=$PWD/chinium_b200/libchinium_fock_$v.so; [ "$v" = base ] && lib=$PWD/chinium_b200/libchinium_fock.so
  for w in ${WORKLOADS:-c18 fe4s4}; do
    CHINIUM_FOCK_LIB=$lib timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_${v}_$w.json 2> gpurun_out/${TAG}_${v}_$w.err
    echo "bench $v $w rc=$?"; head -c 260 gpurun_out/${TAG}_${v}_$w.json | tail -c 120; echo
  done
done
