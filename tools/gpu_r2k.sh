#!/bin/bash
# Round-2 session K: nuclear Hessian (ContractHesss) parity on the GPU; A/B of the interleaved phase-A/B recurrences of the
# warp-group kernels (new vs -DCF_WG_NO_ILP) and of three/four resident CTAs for the classes with few accumulators (alt).
TAG=${TAG:-r2k}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x -k "hess or adaptor_matches or parity or blocks" --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${TAG}_pytest.log
run() {  # name lib variant workload
  CF_WG_VARIANT=$3 CHINIUM_FOCK_LIB=$PWD/chinium_b200/$2 timeout 600 python bench.py --workload $4 --steps 5 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_$1_$4.json 2> gpurun_out/${TAG}_$1_$4.err
  echo "bench $1 $4 rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_$1_$4.json 3
}
for w in c18 fe4s4 h2o64; do
  run new libchinium_fock.so 0 $w
  run noilp libchinium_fock_noilp.so 0 $w
  run alt libchinium_fock_alt.so 1 $w
done
