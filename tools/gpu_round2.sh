#!/bin/bash
# One GPU-box pass of round 2, in stages (one ncu invocation per call):
#   tools/gpu_round2.sh <tag> run     probe of the D pass kernel, gpu tests, smoke, bench (ours, reference, eager)
#   tools/gpu_round2.sh <tag> list    ncu launch list of the bench
#   tools/gpu_round2.sh <tag> full    ncu --set full of the pass kernel, the generator kernels and the GAN-DES GEMM
#   tools/gpu_round2.sh <tag> raster  ncu --set full of the rasteriser kernels
TAG=${1:-run}; STAGE=${2:-run}
mkdir -p gpurun_out
case $STAGE in
run)
  timeout 300 python tools/pass_probe.py all > gpurun_out/probe_$TAG.log 2>&1; rc=$?; echo "probe rc=$rc" >> gpurun_out/probe_$TAG.log
  grep -E "one-kernel|two-kernel|MISMATCH|WATCHDOG" gpurun_out/probe_$TAG.log
  if [ $rc -ne 0 ]; then exit 1; fi
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
  tail -3 gpurun_out/pytest_$TAG.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
  timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 400 gpurun_out/bench_ref_$TAG.json
  timeout 300 python bench.py --impl eager --steps 5 > gpurun_out/bench_eager_$TAG.json 2> gpurun_out/bench_eager_$TAG.err; cat gpurun_out/bench_eager_$TAG.json; tail -2 gpurun_out/bench_eager_$TAG.err
  ;;
list)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu-launch rc=$?"
  ;;
full)
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:disc_pass_fused|gen_layer_tc|gemm_tc_kernel" -s 30 -c 24 -o gpurun_out/full_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu-full rc=$?"
  ;;
raster)
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:raster_" -s 4 -c 2 -o gpurun_out/full_raster_$TAG -f python tools/raster_one.py sort > gpurun_out/ncu_full_raster_$TAG.log 2>&1; echo "ncu-raster rc=$?"
  ;;
esac
ls -la gpurun_out | tail -4
