#!/bin/bash
# Eight-GPU record: cfg4 (8192 pairs, ZoomOut 30->200, oracle-checked sample), cfg5 (599 meshes), cfg2a with e2e.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 400 $RUN 29601 bench.py --gpus 8 --config cfg4 --steps 1 --warmup 1 --verify 8 > gpurun_out/${TAG}_cfg4_8gpu.json 2> gpurun_out/${TAG}_cfg4_8gpu.err; echo "cfg4 rc=$?"
timeout 200 $RUN 29602 bench.py --gpus 8 --config cfg5 > gpurun_out/${TAG}_cfg5_8gpu.json 2> gpurun_out/${TAG}_cfg5_8gpu.err; echo "cfg5 rc=$?"
timeout 200 $RUN 29603 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/${TAG}_cfg2a_8gpu.json 2> gpurun_out/${TAG}_cfg2a_8gpu.err; echo "cfg2a rc=$?"
for f in cfg4 cfg5 cfg2a; do python - gpurun_out/${TAG}_${f}_8gpu.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], d.get("value"), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "bank", (d.get("e2e_bank") or {}).get("value"),
              "verify", d.get("verify"), d.get("gather_transport"), d.get("clocks"))
PY
done
