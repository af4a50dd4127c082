#!/bin/bash
# DRAM traffic + duration of the first N ERI kernels of one build (ncu, 2 metrics): W=<workload> N=<count> RX=<kernel regex>
TAG=${TAG:-d}; W=${W:-h2o64}; N=${N:-12}; RX=${RX:-eri_jk}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k "regex:$RX" -c $N --csv --log-file gpurun_out/${TAG}_dram_$W.csv python bench.py --workload $W --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_dram_$W.log 2>&1
echo "ncu dram rc=$?"
