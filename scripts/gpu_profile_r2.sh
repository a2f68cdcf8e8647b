#!/bin/bash
# Round-2 ncu artefacts: launch list of a short default bench run + full captures of the three kernels that dominate the step
# and of the ZoomOut conversion pass.  Usage: gpu_profile_r2.sh <tag>
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in nn_tc_kernel f2p_tc_kernel fmap_solve32w_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K} -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_${K} \
    python scripts/one_step.py 128 3 > gpurun_out/prof_${TAG}_${K}.log 2>&1
  echo "${K} rc=$?"
done
ls -la gpurun_out | grep prof_${TAG}
