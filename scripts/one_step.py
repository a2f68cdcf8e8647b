"""Two steps of the bench pipeline on a device-resident batch (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from densematcher_b200 import pipeline
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
b = bench.make_host_batch(P).to_device(torch.device("cuda", 0))
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    pipeline.match_pairs_device(b, k=bench.K_EIG, w_descr=bench.W_DESCR, w_lap=bench.W_LAP)
torch.cuda.synchronize()
print("done")
