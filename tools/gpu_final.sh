#!/bin/bash
# Round-end verification session: parity tests, smoke, the driver's bench commands, gradient / multi-density benches, ncu launch list.
TAG=${TAG:-final}
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import json
for f in ("grad_h2o","grad_bo3h3","grad_fe4s4","grad_c18","multi_c18","bench_fe4s4","bench_h2o64","bench_default"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_%s.json"%f).read().strip().splitlines()[-1]); print(f, {k:d.get(k) for k in ("ms_per_step","value","speedup_vs_separate")}, (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f,"ERR",e)
PY
