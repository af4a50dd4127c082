#!/bin/bash
# Round-2 end verification session: parity tests, smoke, the driver's bench commands, gradient / multi-density / Hessian timings,
# ncu launch lists of the bench command (durations, DRAM bytes, RED sectors, FP64 pipe) for c18 and (H2O)64.
TAG=${TAG:-r2z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1; nproc >> gpurun_out/${TAG}_gpu.txt; lscpu | grep 'Model name' >> gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"; head -c 250 gpurun_out/${TAG}_bench_default.json; echo
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>&1; head -c 200 gpurun_out/${TAG}_bench_reference.json; echo
for w in h2o bo3h3; do python bench.py --path grad --workload $w --steps 5 --warmup 2 > gpurun_out/${TAG}_grad_$w.json 2> gpurun_out/${TAG}_grad_$w.err; echo "grad $w rc=$?"; done
for w in fe4s4 c18; do python bench.py --path grad --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_grad_$w.json 2> gpurun_out/${TAG}_grad_$w.err; echo "grad $w rc=$?"; done
python bench.py --path multi --workload c18 --steps 3 --warmup 1 > gpurun_out/${TAG}_multi_c18.json 2> gpurun_out/${TAG}_multi_c18.err; echo "multi rc=$?"
for w in fe4s4 h2o64; do python bench.py --workload $w --steps 3 --warmup 3 --per-class --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "bench $w rc=$?"; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
RX="regex:eri_jk_tpq|eri_jk_wg|pure_to_cart|finalize_kernel|bounds_kernel|scales_kernel"
timeout 900 ncu --metrics $M --clock-control none -k "$RX" -c 420 --csv --log-file gpurun_out/${TAG}_launches_c18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c18.log 2>&1; echo "ncu c18 rc=$?"
timeout 1200 ncu --metrics $M --clock-control none -k "$RX" -c 140 --csv --log-file gpurun_out/${TAG}_launches_h2o64.csv python bench.py --workload h2o64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_h2o64.log 2>&1; echo "ncu h2o64 rc=$?"
timeout 600 python tools/time_hess.py h2o bo3h3 fe4s4 c18 > gpurun_out/${TAG}_hess_timing.jsonl 2> gpurun_out/${TAG}_hess_timing.err; echo "hess timing rc=$?"
python - <<PY
import json
for f in ("grad_h2o","grad_bo3h3","grad_fe4s4","grad_c18","multi_c18","bench_fe4s4","bench_h2o64","bench_default"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1]); print(f, {k:d.get(k) for k in ("ms_per_step","value","speedup_vs_separate")}, (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f,"ERR",e)
PY
