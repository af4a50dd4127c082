#!/bin/bash
# Round-2 session M: full GPU test suite on the new defaults (prim_cut 1e-18, padded staged root tables, recursive-halving
# warp sums for J(a,b)), then A/B against -DCF_NO_WARP_MULTI_SUM and -DCF_TABLE_NOPAD.
TAG=${TAG:-r2m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
run() {  # name lib workload
  CHINIUM_FOCK_LIB=$PWD/chinium_b200/$2 timeout 600 python bench.py --workload $3 --steps 5 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_$1_$3.json 2> gpurun_out/${TAG}_$1_$3.err
  echo "bench $1 $3 rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_$1_$3.json 3
}
for w in c18 fe4s4 h2o64; do
  run new libchinium_fock.so $w
  run nowms libchinium_fock_nowms.so $w
  run nopad libchinium_fock_nopad.so $w
done
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for v in new nopad; do
  lib=libchinium_fock.so; [ $v = nopad ] && lib=libchinium_fock_nopad.so
  CHINIUM_FOCK_LIB=$PWD/chinium_b200/$lib timeout 600 ncu --metrics $M --clock-control none -k "regex:eri_jk_tpq|eri_jk_wg" -c 61 --csv --log-file gpurun_out/${TAG}_smem_${v}_c18.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_$v.log 2>&1; echo "ncu $v rc=$?"
done
