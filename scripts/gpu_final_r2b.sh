#!/bin/bash
# Final one-GPU record after the bank kernels / launch merges: tests, smoke, default bench line, reference arm, the other
# configurations (CPU legs of those are in profiles/bench_r2_final_*), ncu launch list of the default bench command.
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_cfg2a.json 2> gpurun_out/${TAG}_cfg2a.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"
for c in cfg2b cfg3; do
  timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err; echo "$c rc=$?"
done
timeout 900 python bench.py --config cfg4 --pairs 512 --steps 1 --warmup 1 --verify 2 --no-cpu-baseline > gpurun_out/${TAG}_cfg4.json 2> gpurun_out/${TAG}_cfg4.err; echo "cfg4 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu rc=$?"
for f in cfg2a ref cfg2b cfg3 cfg4; do python - gpurun_out/${TAG}_$f.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], d.get("impl", "ours"), d.get("value"), d.get("unit"), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"),
              "bank", (d.get("e2e_bank") or {}).get("value"), "verify", d.get("verify"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
done
