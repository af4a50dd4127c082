#!/bin/bash
# Round-2 session H: partition tests after the task-ownership change, ncu launch list of the bench command (duration,
# DRAM bytes, RED sectors, warp instructions per launch) for c18 and (H2O)64.
TAG=${TAG:-r2h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "partition or two_limb or golden or jk_parity" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c18.log 2>&1; echo "ncu c18 rc=$?"
timeout 1200 ncu --metrics $M --clock-control none -c 160 --csv --log-file gpurun_out/${TAG}_launches_h2o64.csv python bench.py --workload h2o64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_h2o64.log 2>&1; echo "ncu h2o64 rc=$?"
ls -la gpurun_out/${TAG}_launches_*.csv
