#!/bin/bash
# ncu launch list (per-launch durations) of the LAST of three pipeline steps
set -u
TAG=${1:-r1}; P=${2:-128}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
  python scripts/one_step.py $P 3 > gpurun_out/launches_${TAG}.log 2>&1
echo "rc=$?"
