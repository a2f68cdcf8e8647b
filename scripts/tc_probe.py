"""Quick on-GPU probe of the tcgen05 engine: score error vs float64 and a small argmax, with prints after each step
(run under `timeout`: a hang then shows which step did not return)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200 import nn, _lib
print(_lib.load().dm_build_info().decode(), flush=True)
rng = np.random.default_rng(0)
for (nq, nd, d) in [(128, 256, 64), (128, 256, 384), (300, 700, 100), (2000, 2000, 384)]:
    Y = rng.standard_normal((nq, d)).astype(np.float32); X = rng.standard_normal((nd, d)).astype(np.float32)
    t0 = time.time()
    S = nn.debug_scores(torch.from_numpy(Y).cuda(), torch.from_numpy(X).cuda()); torch.cuda.synchronize()
    S = S.cpu().numpy().astype(np.float64)
    S64 = Y.astype(np.float64) @ X.astype(np.float64).T
    nrm = np.linalg.norm(Y, axis=1)[:, None] * np.linalg.norm(X, axis=1)[None, :]
    err = np.abs(S - S64) / nrm
    print(f"scores nq={nq} nd={nd} d={d}: max err/(|y||x|) = {err.max():.3e}  mean {err.mean():.3e}  ({time.time()-t0:.2f}s)", flush=True)
    if err.max() > 1e-3:
        bad = np.argwhere(err > 1e-3)
        print("  BAD entries:", len(bad), "first:", bad[:5].tolist(), "rows bad:", np.unique(bad[:, 0])[:10], "cols bad:", np.unique(bad[:, 1])[:10], flush=True)
for (nq, nd, d) in [(130, 200, 64), (2000, 2000, 384)]:
    Y = rng.standard_normal((nq, d)).astype(np.float32); X = rng.standard_normal((nd, d)).astype(np.float32)
    (r,), (c,), st = nn.nn_argmax(torch.from_numpy(Y).cuda(), torch.from_numpy(X).cuda(), col_epi=(nn.COSINE_UNIT,), return_stats=True)
    S64 = Y.astype(np.float64) @ X.astype(np.float64).T
    print(f"argmax nq={nq} nd={nd} d={d}: row mismatches {int((r.cpu().numpy() != S64.argmax(1)).sum())} col mismatches {int((c.cpu().numpy() != S64.argmax(0)).sum())} stats {st}", flush=True)
print("probe done", flush=True)
