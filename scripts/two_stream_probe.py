"""Two independent batches in flight on two streams: does the HBM-bound operand preparation of one overlap the
tensor / ALU-bound passes of the other?  python scripts/two_stream_probe.py [pairs per batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from densematcher_b200 import pipeline, nn as dnn
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
batches = [bench.make_host_batch(P, seed=2000 + i).to_device(dev) for i in range(2)]
kw = dict(k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP, check=False)
def timed(fn, n_pairs, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, n_pairs / ms * 1e3
ms, r = timed(lambda: pipeline.match_pairs_device(batches[0], **kw), P)
print(f"one stream, {P} pairs per call: {ms:.3f} ms -> {r:.0f} pairs/s")
streams = [torch.cuda.Stream(dev) for _ in range(2)]
def two():
    cur = torch.cuda.current_stream(dev)
    for s, b in zip(streams, batches):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            pipeline.match_pairs_device(b, **kw)
    for s in streams: cur.wait_stream(s)
ms, r = timed(two, 2 * P)
print(f"two streams, 2 x {P} pairs in flight: {ms:.3f} ms -> {r:.0f} pairs/s")
