#!/bin/bash
# full ncu capture of the functional-map kernels (fp64 GEMM, Cholesky solve) of the second pipeline step;
# the raw page is exported to CSV on the box (the .ncu-rep itself is only kept when small).
set -u
TAG=${1:-r1_fm}; KREG=${2:-"gemm64_kernel|fmap_solve_kernel"}; SKIP=${3:-7}; CNT=${4:-7}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREG}" -s ${SKIP} -c ${CNT} -f -o gpurun_out/prof_${TAG} \
  python scripts/one_step.py 64 > gpurun_out/prof_${TAG}.log 2>&1
echo "capture rc=$?"
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
sz=$(stat -c %s gpurun_out/prof_${TAG}.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/prof_${TAG}.ncu-rep; fi
