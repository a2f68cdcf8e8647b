"""Where the bank-shaped host entry spends a 128-pair call: chunk size sweep, pieces timed separately."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline
from densematcher_b200.pipeline import MeshBankHost
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
host = bench.make_host_batch(P)
pool = host.pool
N = bench.N_VERT
hbank = MeshBankHost(F=np.concatenate(pool["feats"]), off=np.arange(9, dtype=np.int64) * N,
                     Phi=np.concatenate([b[1] for b in pool["bases"]]), evals=np.stack([b[0] for b in pool["bases"]]),
                     area=np.concatenate([b[2] for b in pool["bases"]])).pin()
kw = dict(k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP, copy=False)
def tm(f, n=5):
    f(); f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
for ch in (16, 32, 64, 128):
    ms = tm(lambda: pipeline.match_bank_pairs_host(hbank, pool["ia"], pool["ib"], dev, chunk_pairs=ch, **kw))
    print(f"chunk={ch}: {ms:.2f} ms per {P} pairs = {P / ms * 1e3:.0f} pairs/s", flush=True)
# pieces at chunk = 128
dbank = pipeline.MeshBankDevice(hbank.F, hbank.off, hbank.Phi, hbank.evals, hbank.area, device=dev)
try:
    b = dbank.assemble(pool["ia"], pool["ib"])
    print(f"assemble: {tm(lambda: dbank.assemble(pool['ia'], pool['ib'])):.2f} ms")
    print(f"match_pairs_device: {tm(lambda: pipeline.match_pairs_device(b, check=False, k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP)):.2f} ms")
except Exception as e:
    print("pieces failed:", repr(e))
# the bank kernels: pairs as id lists, nothing assembled (dm_bank_prepare once + dm_match_bank_pairs)
kwd = dict(k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP, check=False)
dbank.prepared(bench.K_EIG)
print(f"bank.match (prepared): {tm(lambda: dbank.match(pool['ia'], pool['ib'], **kwd)):.2f} ms per {P} pairs")
def prep():
    dbank._states.clear()
    dbank.prepared(bench.K_EIG)
print(f"bank prepare (8 meshes): {tm(prep):.2f} ms")
if os.environ.get("PROBE_ONE"):  # one more call for an ncu launch list (skip the warm-up launches with --launch-skip)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    dbank.match(pool['ia'], pool['ib'], **kwd); torch.cuda.synchronize(); torch.cuda.profiler.stop()
