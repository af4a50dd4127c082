#!/bin/bash
# multi-GPU scaling session on one box: usage NS="4 8" WORKLOADS="c18 h2o64" tools/gpu_scale.sh
TAG=${TAG:-scale}
mkdir -p gpurun_out
for w in ${WORKLOADS:-c18}; do
  for n in ${NS:-8}; do
    st=5; [ "$w" = h2o64 ] && st=3
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps $st --warmup 3 --workload $w > gpurun_out/${TAG}_n${n}_$w.json 2> gpurun_out/${TAG}_n${n}_$w.err
    echo "n=$n $w rc=$?"; python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_n${n}_$w.json").read().strip().splitlines()[-1]); print({k:d[k] for k in ("n_gpus","ms_per_step","value")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])
except Exception as e: print("ERR",e)
EOF
  done
done
