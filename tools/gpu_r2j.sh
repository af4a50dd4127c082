#!/bin/bash
# Round-2 session J: shared-memory strides of the warp-group kernels from the bank model (tools/wg_bank_model.py).
# A/B of three builds (new strides / old strides / alternative configurations) + ncu shared-memory counters of the wg kernels.
TAG=${TAG:-r2j}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
CHINIUM_FOCK_LIB=$PWD/chinium_b200/libchinium_fock.so timeout 900 python -m pytest tests -m gpu -q -x -k "parity or blocks" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
run() {  # name lib variant workload
  CF_WG_VARIANT=$3 CHINIUM_FOCK_LIB=$PWD/chinium_b200/$2 timeout 600 python bench.py --workload $4 --steps 5 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_$1_$4.json 2> gpurun_out/${TAG}_$1_$4.err
  echo "bench $1 $4 rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_$1_$4.json 3
}
for w in c18 fe4s4 h2o64; do
  run new libchinium_fock.so 0 $w
  run old libchinium_fock_old.so 0 $w
  run alt libchinium_fock_alt.so 1 $w
done
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for v in new old; do
  lib=libchinium_fock.so; [ $v = old ] && lib=libchinium_fock_old.so
  CHINIUM_FOCK_LIB=$PWD/chinium_b200/$lib timeout 600 ncu --metrics $M --clock-control none -k "regex:eri_jk_wg" -c 40 --csv --log-file gpurun_out/${TAG}_smem_${v}_c18.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$v.log 2>&1; echo "ncu $v rc=$?"
done
