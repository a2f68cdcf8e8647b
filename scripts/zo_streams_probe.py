"""Does running two half-batches of the ZoomOut ladder on two streams hide the small kernels of one half behind the
score passes of the other?  python scripts/zo_streams_probe.py [pairs] [splits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from densematcher_b200 import fm as dfm, nn as dnn, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
rng = np.random.default_rng(4000)
n, K, nit = 2000, 200, 170
pool = [synth.synthetic_basis(n, K, rng) for _ in range(4)]
ia, ib = rng.integers(0, 4, P), rng.integers(0, 4, P)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
Phi1, Phi2 = up(np.concatenate([pool[i][1] for i in ia])), up(np.concatenate([pool[i][1] for i in ib]))
a2 = up(np.concatenate([pool[i][2] for i in ib]))
C0 = up(np.stack([np.linalg.qr(rng.standard_normal((30, 30)))[0] for _ in range(P)]))

def run(splits):
    streams = [torch.cuda.Stream(dev) for _ in range(splits)]
    wss = [dnn.Workspace(dev) for _ in range(splits)]
    per = P // splits
    offs = []
    for s in range(splits):
        o = np.arange(per + 1) * n
        offs.append(dnn.Offsets(torch.from_numpy(o).to(dev), o))
    def once():
        cur = torch.cuda.current_stream(dev)
        outs = []
        for s in range(splits):
            streams[s].wait_stream(cur)
            with torch.cuda.stream(streams[s]):
                r0, r1 = s * per * n, (s + 1) * per * n
                outs.append(dfm.zoomout(C0[s * per:(s + 1) * per], Phi1[r0:r1], Phi2[r0:r1], a2[r0:r1], nit, 1, offs[s], offs[s],
                                        return_p2p=True, out_dtype=torch.int32, workspace=wss[s]))
        for s in range(splits):
            cur.wait_stream(streams[s])
        return outs
    once(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2): outs = once()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print(f"{splits} stream(s): {ms:.1f} ms per {P} pairs = {P / ms * 1e3:.0f} pairs/s")
    return outs

ref = run(1)
for sp in (2, 4):
    got = run(sp)
    p_ref = ref[0][1]
    p_got = torch.cat([g[1] for g in got])
    print("   same final p2p:", bool(torch.equal(p_ref, p_got)))
