#!/bin/bash
# One-GPU record of a round: smoke, the default bench line, the reference arm, the other BASELINE configurations.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_cfg2a.json 2> gpurun_out/${TAG}_cfg2a.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"
for c in cfg2b cfg3 cfg5; do
  timeout 600 python bench.py --config $c > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err; echo "$c rc=$?"
done
timeout 900 python bench.py --config cfg4 --pairs 512 --steps 1 --warmup 1 --verify 2 > gpurun_out/${TAG}_cfg4.json 2> gpurun_out/${TAG}_cfg4.err; echo "cfg4 rc=$?"
for f in cfg2a ref cfg2b cfg3 cfg5 cfg4; do python - gpurun_out/${TAG}_$f.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], d.get("impl", "ours"), d.get("value"), d.get("unit"), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"),
              "bank", (d.get("e2e_bank") or {}).get("value"), "verify", d.get("verify"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
done
