#!/bin/bash
# usage (on the GPU box): tools/gpu_dp_bench.sh <tag> <N> : bench at N GPUs with per-replica and SyncBN generator statistics
TAG=$1; N=$2; O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/dp_${TAG}_gpus.txt 2>&1
for extra in "" "--sync-bn"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-raster $extra \
     > $O/bench_dp${N}_${TAG}${extra}.json 2> $O/bench_dp${N}_${TAG}${extra}.err; echo "bench $extra rc=$?"
  tail -c 600 $O/bench_dp${N}_${TAG}${extra}.json
done
