#!/bin/bash
# ncu launch list of a short bench run + one full capture of the named kernel.  Usage: gpu_profile.sh <tag> <kernel-regex>
set -u
TAG=${1:-r1}; KREG=${2:-nn_ffma_kernel}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --pairs 64 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREG} -s 3 -c 2 -f -o gpurun_out/prof_${TAG} \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --pairs 64 > gpurun_out/prof_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
