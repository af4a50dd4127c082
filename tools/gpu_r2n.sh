#!/bin/bash
# Round-2 session N (2 GPUs): the bench contract under torchrun at N = 2 (both arms), the multi-device tests.
TAG=${TAG:-r2n}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
for w in c18 h2o64; do
  st=5; [ "$w" = h2o64 ] && st=3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps $st --warmup 3 --workload $w > gpurun_out/${TAG}_n2_$w.json 2> gpurun_out/${TAG}_n2_$w.err
  echo "n=2 $w rc=$?"; tail -c 600 gpurun_out/${TAG}_n2_$w.json | head -c 300; echo
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/${TAG}_n2_reference.json 2> gpurun_out/${TAG}_n2_reference.err; echo "reference arm rc=$?"; head -c 300 gpurun_out/${TAG}_n2_reference.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/dist_grad_check.py bo3h3 > gpurun_out/${TAG}_dist_check.log 2>&1; echo "dist check rc=$?"; tail -3 gpurun_out/${TAG}_dist_check.log
