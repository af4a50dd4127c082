#!/bin/bash
# A/B session: quick parity subset + benches with per-class table.  TAG=<name> WORKLOADS="c18 fe4s4 h2o64" TESTS="<pytest -k expr>"
TAG=${TAG:-ab}
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -x -k "$TESTS" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
fi
for w in ${WORKLOADS:-c18}; do
  timeout 900 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --per-class --no-cpu-baseline $BENCH_FLAGS > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_$w.json 8; tail -3 gpurun_out/${TAG}_bench_$w.err
done
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
