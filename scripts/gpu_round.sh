#!/bin/bash
# One GPU-box visit: smoke, parity tests, a short bench, the ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
ls /root/reference > gpurun_out/ref_present.txt 2>&1
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
