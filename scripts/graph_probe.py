"""Single-pair latency: eager launches vs CUDA-graph replay of the same stream-ordered C-ABI call sequence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from densematcher_b200 import pipeline, fm as dfm, nn as dnn
dev = torch.device("cuda", 0)
def tm(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
for P in (1, 8):
    b = bench.make_host_batch(P).to_device(dev)
    kw = dict(k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP, out_dtype=torch.int32)
    step = lambda: pipeline.match_pairs_device(b, check=False, **kw)
    ref = step(); torch.cuda.synchronize()
    print(f"P={P}: eager step {tm(step):.3f} ms", flush=True)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): step()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = step()
        torch.cuda.synchronize()
        print(f"P={P}: graph replay {tm(g.replay):.3f} ms; equal to eager: "
              f"{all(torch.equal(out[n], ref[n]) for n in ref if n != 'C')} C close {float((out['C']-ref['C']).abs().max()):.1e}", flush=True)
    except Exception as e:
        print("graph capture failed:", repr(e)[:300], flush=True)
# ZoomOut 30->100, single pair
rng = np.random.default_rng(0)
from oracle import meshgen
e1, P1, a1 = meshgen.synthetic_basis(2000, 100, rng); e2, P2, a2 = meshgen.synthetic_basis(2000, 100, rng)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
C0 = up(np.linalg.qr(rng.standard_normal((30, 30)))[0][None]); P1d, P2d, a2d = up(P1), up(P2), up(a2)
off = dnn.Offsets(torch.tensor([0, 2000], device=dev), np.array([0, 2000]))
zo = lambda: dfm.zoomout(C0, P1d, P2d, a2d, 70, 1, off, off, return_p2p=True, out_dtype=torch.int32)
r = zo(); torch.cuda.synchronize()
print(f"zoomout 30->100 single pair: eager {tm(zo, 5):.2f} ms", flush=True)
try:
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        zo()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        o = zo()
    torch.cuda.synchronize()
    print(f"zoomout graph replay {tm(g.replay, 5):.2f} ms; equal: {torch.equal(o[1], r[1])} {float((o[0]-r[0]).abs().max()):.1e}", flush=True)
except Exception as e:
    print("zoomout graph capture failed:", repr(e)[:300], flush=True)
