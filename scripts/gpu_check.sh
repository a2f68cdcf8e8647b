#!/bin/bash
# First-line GPU check: build info, the GPU test-suite under a timeout (a hung kernel must not hang the box), smoke.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/nvidia_smi.txt 2>&1
timeout ${T1:-900} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
