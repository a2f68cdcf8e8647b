#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (bench.py measures configs[1]); one JSON line each.

    python scripts/bench_configs.py [cfg3] [cfg4] [cfg4fast] [cfg5] [--pairs-scale S]

cfg3: batch of 1024 ragged pairs (N ~ U(1500, 2500), d = 384), feature NN only, both directions.
cfg4: ZoomOut ladder k = 30 -> 200, step 1, N = 2000 (the per-GPU share of 8192 pairs is 1024; a 128-pair batch is timed).
cfg5: DenseCorr3D stand-in: 599 meshes in 24 categories (N ~ U(1800, 2200), d = 384, K = 100), every ordered
      intra-category pair, from a device-resident mesh bank (full hot path per pair).
Under torchrun every rank processes its block of pairs (weak scaling for cfg3 / cfg4, strong for cfg5).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from densematcher_b200 import nn as dnn, fm as dfm, pipeline, _lib
from oracle import meshgen

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, steps=3, warm=1):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(name, pairs_total, ms, extra):
    if rank == 0:
        print(json.dumps({"config": name, "metric": "mesh_pairs_per_sec", "value": pairs_total / (ms * 1e-3), "unit": "pairs/s",
                          "n_gpus": world, "ms_per_step": ms, "pairs_per_step": pairs_total, "data": "synthetic", **extra}))


def cfg3():
    g = torch.Generator(device=dev).manual_seed(3000 + rank)
    rng = np.random.default_rng(3000 + rank)
    P = 1024
    nq, nd = rng.integers(1500, 2501, size=P), rng.integers(1500, 2501, size=P)
    qo, do = np.concatenate([[0], np.cumsum(nq)]), np.concatenate([[0], np.cumsum(nd)])
    Y = torch.nn.functional.normalize(torch.randn(int(qo[-1]), 384, device=dev, generator=g), dim=1)
    X = torch.nn.functional.normalize(torch.randn(int(do[-1]), 384, device=dev, generator=g), dim=1)
    qoff, doff = dnn.Offsets(torch.from_numpy(qo).to(dev), qo), dnn.Offsets(torch.from_numpy(do).to(dev), do)
    fn = lambda: dnn.nn_argmax(Y, X, qoff, doff, row_epi=(dnn.COSINE_UNIT,), col_epi=(dnn.COSINE_UNIT,), out_dtype=torch.int32)
    ms = timed(fn)
    flops = 2.0 * float(np.sum(nq.astype(np.float64) * nd)) * 384
    emit("cfg3: 1024 ragged pairs/GPU, N~U(1500,2500), d=384, NN both directions", P * world, ms,
         {"nn_stage_algorithmic_tflops_per_gpu": flops / (ms * 1e-3) / 1e12})


def cfg4(fast):
    rng = np.random.default_rng(4000 + rank)
    P, n, K, nit = 128, 2000, 200, 170
    pool = [meshgen.synthetic_basis(n, K, rng) for _ in range(4)]
    ia, ib = rng.integers(0, 4, P), rng.integers(0, 4, P)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Phi1, Phi2 = up(np.concatenate([pool[i][1] for i in ia])), up(np.concatenate([pool[i][1] for i in ib]))
    a2 = up(np.concatenate([pool[i][2] for i in ib]))
    off = np.arange(P + 1) * n
    o = dnn.Offsets(torch.from_numpy(off).to(dev), off)
    C0 = up(np.stack([np.linalg.qr(rng.standard_normal((30, 30)))[0] for _ in range(P)]))
    fl = _lib.DM_FAST_FM if fast else 0
    fn = lambda: dfm.zoomout(C0, Phi1, Phi2, a2, nit, 1, o, o, return_p2p=True, out_dtype=torch.int32, flags=fl)
    ms = timed(fn, steps=2)
    emit("cfg4: ZoomOut ladder k=30->200 step 1, N=2000, 128-pair batch/GPU" + (" (DM_FAST_FM)" if fast else " (float64 C)"),
         P * world, ms, {"seconds_for_8192_pairs_on_this_many_gpus": 8192 / (P * world / (ms * 1e-3))})


def cfg5():
    rng = np.random.default_rng(5000)
    n_meshes, n_cat, d, K = 599, 24, 384, 100
    cats = np.sort(rng.integers(0, n_cat, size=n_meshes))
    sizes = rng.integers(1800, 2201, size=n_meshes)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    g = torch.Generator(device=dev).manual_seed(5000)
    F = torch.nn.functional.normalize(torch.randn(int(off[-1]), d, device=dev, generator=g), dim=1)
    # synthetic A-orthonormal bases generated on the device (QR per mesh would take minutes on the host)
    area = (torch.rand(int(off[-1]), device=dev, generator=g, dtype=torch.float64) + 0.5) / 2000.0
    Phi = torch.empty(int(off[-1]), K, device=dev, dtype=torch.float64)
    for i in range(n_meshes):
        s = slice(off[i], off[i + 1])
        M = torch.randn(int(sizes[i]), K, device=dev, generator=g, dtype=torch.float64)
        M[:, 0] = 1.0
        Q, R = torch.linalg.qr(torch.sqrt(area[s])[:, None] * M)
        Phi[s] = Q * torch.sign(torch.diagonal(R))[None, :] / torch.sqrt(area[s])[:, None]
    evals = torch.cumsum(torch.rand(n_meshes, K, device=dev, generator=g, dtype=torch.float64), dim=1)
    evals[:, 0] = 0.0
    bank = pipeline.MeshBankDevice(F, off, Phi=Phi, evals=evals, area=area, device=dev)
    src, dst = pipeline.intra_category_pairs(cats)
    kw = dict(k=K, w_descr=1e4, w_lap=1e3, out_dtype=torch.int32)
    fn = lambda: pipeline.match_bank_pairs(bank, src, dst, chunk_pairs=128, rank=rank, world=world, to_host=False, **kw)
    ms = timed(fn, steps=1, warm=1)
    emit(f"cfg5 stand-in: 599 meshes / 24 categories, all {len(src)} ordered intra-category pairs from a device mesh bank, "
         "full hot path (NN + projection + solve + FM->p2p)", len(src), ms, {"scaling": "strong"})


def extras():
    """SURVEY 8f rank 2 on a batch of 64 copies of a real mesh pair (two deformations of icosphere(4): 2562 vertices,
    5120 faces, cotangent LBO basis, k = 50): Hungarian assignment of every pair's mapped indicator and the barycentric
    precise map."""
    n_pairs, k = 64, 50
    V, F = meshgen.icosphere(4)
    V1 = meshgen.deform(V, (1.0, 0.9, 1.1), 0.1, (0.3, 0.2))
    V2 = meshgen.deform(V, (1.1, 1.0, 0.85), 0.15, (1.0, 0.5))
    _, P1, a1 = meshgen.lbo_basis(V1, F, k)
    _, P2, a2 = meshgen.lbo_basis(V2, F, k)
    n, nf = len(V), len(F)
    C1 = P2.T @ (a2[:, None] * P1)                                     # functional map of the identity vertex map
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    off = np.arange(n_pairs + 1, dtype=np.int64) * n
    batch = pipeline.PairBatchDevice(None, None, off, off, dev, Phi1=t(P1).repeat(n_pairs, 1), Phi2=t(P2).repeat(n_pairs, 1),
                                     area1=t(a1).repeat(n_pairs), area2=t(a2).repeat(n_pairs))
    rng = np.random.default_rng(3)
    C = torch.stack([t(C1 + 1e-3 * rng.standard_normal(C1.shape)) for _ in range(n_pairs)]).contiguous()
    ms = timed(lambda: pipeline.hungarian_pairs(batch, C, chunk_pairs=64), steps=1, warm=1)
    emit(f"extras: Hungarian assignment (dm_lap_solve) of the {n} x {n} mapped indicator, icosphere(4) pair, k = {k}, "
         f"{n_pairs} pairs (scipy: ~0.3 s per pair and core)", n_pairs, ms, {})
    faces = t(F, torch.int32).repeat(n_pairs, 1)
    foff = np.arange(n_pairs + 1, dtype=np.int64) * nf
    ms = timed(lambda: pipeline.precise_maps(batch, C, faces, foff), steps=3, warm=1)
    emit(f"extras: barycentric precise map (dm_precise_map), {n} points on {nf} faces, p = {k}, {n_pairs} pairs", n_pairs, ms, {})


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["cfg3", "cfg4", "cfg4fast", "cfg5", "extras"]
    for w in which:
        {"cfg3": cfg3, "cfg4": lambda: cfg4(False), "cfg4fast": lambda: cfg4(True), "cfg5": cfg5, "extras": extras}[w]()
    if world > 1:
        dist.destroy_process_group()
