#!/bin/bash
# usage (on the GPU box): tools/gpu_dp.sh <tag> <N> : NCCL parity test (N >= 2) + bench at N GPUs with per-replica and SyncBN generator statistics
TAG=$1; N=$2; O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/dp_${TAG}_gpus.txt 2>&1
timeout 500 python -m pytest tests/test_gpu_dp_nccl.py tests/test_gpu_trainer.py -q -m gpu -x > $O/pytest_dp_${TAG}.log 2>&1; echo "pytest rc=$?" >> $O/pytest_dp_${TAG}.log
tail -5 $O/pytest_dp_${TAG}.log
for extra in "" "--sync-bn"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 $extra \
     > $O/bench_dp${N}_${TAG}${extra}.json 2> $O/bench_dp${N}_${TAG}${extra}.err; echo "bench $extra rc=$?"
  tail -c 1500 $O/bench_dp${N}_${TAG}${extra}.json
done
