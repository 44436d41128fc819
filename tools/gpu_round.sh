#!/bin/bash
# One GPU-box pass: gpu tests, smoke, bench (ours + reference arm), ncu launch list, ncu full capture of one iteration's tcgen05 kernels.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu-launch rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:disc_bwd_fused|disc_fwd_fused|gen_layer_tc" -s 52 -c 26 -o gpurun_out/full_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-raster > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu-full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:raster_" -s 4 -c 2 -o gpurun_out/full_raster_$TAG -f python tools/raster_one.py sort > gpurun_out/ncu_full_raster_$TAG.log 2>&1; echo "ncu-raster rc=$?"
ls -la gpurun_out | tail -8
