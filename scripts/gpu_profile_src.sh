#!/bin/bash
# full ncu capture of ONE launch of a kernel of the bench step + source-page export (SASS with stall samples)
set -u
TAG=${1:-r1_src}; KREG=${2:-"nn_tc_kernel<1, 1"}; SKIP=${3:-2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREG}" -s ${SKIP} -c 1 -f -o gpurun_out/prof_${TAG} \
  python scripts/one_step.py 64 3 > gpurun_out/prof_${TAG}.log 2>&1
echo "capture rc=$?"
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
rm -f gpurun_out/prof_${TAG}.ncu-rep
