#!/bin/bash
# Round-2 session I: gradient matrices + partition tests, ncu launch lists of the BUILD kernels of the bench command.
TAG=${TAG:-r2i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "matrix_form or partition or contract_grads or cpp_adaptor or multi_density" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
RX="regex:eri_jk_tpq|eri_jk_wg|pure_to_cart|finalize_kernel|bounds_kernel|scales_kernel"
timeout 900 ncu --metrics $M --clock-control none -k "$RX" -c 420 --csv --log-file gpurun_out/${TAG}_launches_c18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c18.log 2>&1; echo "ncu c18 rc=$?"
timeout 1200 ncu --metrics $M --clock-control none -k "$RX" -c 140 --csv --log-file gpurun_out/${TAG}_launches_h2o64.csv python bench.py --workload h2o64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_h2o64.log 2>&1; echo "ncu h2o64 rc=$?"
ls -la gpurun_out/${TAG}_launches_*.csv
