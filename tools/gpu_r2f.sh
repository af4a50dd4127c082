#!/bin/bash
# Round-2 session F: full parity suites after the device-side setup, one-electron integrals, device-resident SCF, benches
TAG=${TAG:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_baseline.py tests/test_gpu.py -m gpu -q -x --durations=6 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${TAG}_pytest.log
for w in h2o64 c18 fe4s4; do
  CF_SETUP_TIMING=1 timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --per-class > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
  echo "bench $w rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench_$w.json 6; grep cf_create gpurun_out/${TAG}_bench_$w.err | head -8
done
timeout 600 python tools/bench_scf_device.py h2o64 4 > gpurun_out/${TAG}_scf_device_h2o64.json 2> gpurun_out/${TAG}_scf_device_h2o64.err; cat gpurun_out/${TAG}_scf_device_h2o64.json; tail -2 gpurun_out/${TAG}_scf_device_h2o64.err
timeout 300 python tools/bench_scf_device.py c18 4 > gpurun_out/${TAG}_scf_device_c18.json 2>/dev/null; cat gpurun_out/${TAG}_scf_device_c18.json
