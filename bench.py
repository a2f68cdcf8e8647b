#!/usr/bin/env python
"""Benchmarks of the correspondence hot path on B200: mesh pairs matched per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2a|cfg2b|cfg3|cfg4|cfg5] [--pairs B]
                    [--impl ours|reference] [--verify V]

Default (the headline, BASELINE.json configs[1] replicated): cfg2a -- one "step" = one pass of the hot path over a batch
of B = 128 synthetic mesh pairs per GPU (N = M = 2000 vertices, d = 384 unit features, k = 100 LBO basis; 9.4 MB of input
per pair, so a step's inputs are ~10x the L2):
    feature NN (cosine argmax, both directions, one fused pass) -> projection Phi^T A F (both meshes) ->
    closed-form C (k = 100) -> FM->p2p (kd-tree-equivalent pair + dense-argmax pair from one pass).
The other BASELINE.json configurations print the same JSON schema:
    cfg2b  cfg2a + ZoomOut 30 -> 100 (70 rungs) + final p2p                       (configs[1] with its ZoomOut)
    cfg3   1024 ragged pairs per GPU, N ~ U(1500, 2500), NN only                   (configs[2])
    cfg4   8192 pairs in total, C0 = B A^+ at k = 30, ZoomOut ladder 30 -> 200, sharded over the ranks   (configs[3])
    cfg5   DenseCorr3D stand-in: 599 meshes / 24 categories, every ordered intra-category pair, sharded   (configs[4])
Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: the same through the host-buffer entry (pinned
H2D of every input + D2H of every result inside the timed region).  `--impl reference` times the CPU restatement of the
reference path (oracle/: the reference's own sklearn kd-tree call + numpy/scipy float64) on a bounded sample.
Multi-GPU: pairs shard across ranks with no data-path collective; every step ends with the NCCL all-gather of all
index maps and C.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VERT, D_FEAT, K_EIG = 2000, 384, 100
W_DESCR, W_LAP = 1e4, 1e3
METRIC = "mesh_pairs_per_sec"
UNIT = "pairs/s"
ALG_BYTES_NN = 4 * D_FEAT * 2 * N_VERT + 4 * 2 * N_VERT          # 6.160 MB per pair
ALG_FLOPS_NN = 2 * N_VERT * N_VERT * D_FEAT                      # 3.072 GFLOP per pair
NCU_DRAM_BYTES_PER_PAIR = (794.701312e6 + 78.630656e6) / 128       # nn_tc_kernel<1,1>: profiles/r2_nn_tc_full_raw.csv (128 pairs)
NCU_F2P_DRAM_BYTES_PER_PAIR = (272.526080e6 + 136.424704e6) / 128   # f2p_tc_kernel<2>: profiles/r2_f2p_tc_full_raw.csv
NCU_SOLVE_DRAM_BYTES_PER_PAIR = (23.4176e6 + 0.112128e6) / 128       # fmap_solve32w_kernel<4>: profiles/r2_fmap_solve32w_full_raw.csv
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12                # CUDA-core FFMA peak at max clock
LANE_OPS_PER_S = 148 * 128 * 1.965e9                             # issue-limited lane instructions per second (4 x 32 lanes per SM)
ALU_PIPE_OPS_PER_S = 148 * 64 * 1.965e9                          # the ALU pipe (min / max / compare / logic) runs at half rate: 16 lanes / clk / scheduler

CONFIG_DEFAULTS = {  # pairs (per GPU for weak configs, total for strong ones), default steps / warmup when not given
    "cfg2a": dict(pairs=128, scaling="weak"), "cfg2b": dict(pairs=64, scaling="weak"),
    "cfg3": dict(pairs=1024, scaling="weak"), "cfg4": dict(pairs=8192, scaling="strong"),
    "cfg5": dict(pairs=None, scaling="strong"),
}


def workload_config(cfg, pairs, n_gpus):
    text = {
        "cfg2a": f"cfg2a x{pairs}/GPU: pairs of N=M={N_VERT} meshes, d={D_FEAT} unit features, k={K_EIG} LBO basis; "
                 "feature NN (both directions) + projection + closed-form C + FM->p2p (4 index maps)",
        "cfg2b": f"cfg2b x{pairs}/GPU: cfg2a + ZoomOut ladder k=30->100 step 1 (70 rungs, upstream pyFM semantics) from "
                 "C[:30,:30] + final p2p",
        "cfg3": f"cfg3: {pairs} ragged pairs/GPU, N~U(1500,2500), d={D_FEAT}, feature NN both directions",
        "cfg4": f"cfg4: {pairs} pairs in total (N={N_VERT}, d={D_FEAT}), C0 = B A^+ at k=30 from the feature projections, "
                "ZoomOut ladder k=30->200 step 1 (170 rungs) + final p2p, pairs sharded over the ranks",
        "cfg5": "cfg5 stand-in (DenseCorr3D is not in the container): 599 meshes / 24 categories, N~U(1800,2200), "
                f"d={D_FEAT}, K={K_EIG}; every ordered intra-category pair ({pairs}) from a device mesh bank, full hot path "
                "(NN + projection + solve + FM->p2p); every step = the whole job: once-per-mesh preparation of the bank "
                "(dm_bank_prepare) + all pairs as id lists (dm_match_bank_pairs), pairs sharded over the ranks",
    }[cfg]
    per_pair_mb = {"cfg2a": 9.4, "cfg2b": 9.4, "cfg3": 6.2, "cfg4": 12.5, "cfg5": 9.4}[cfg]
    return {"workload": text, "name": cfg, "pairs": pairs, "n": N_VERT, "d": D_FEAT, "k": K_EIG, "w_descr": W_DESCR,
            "w_lap": W_LAP,
            "parallelism": f"pairs sharded over {n_gpus} rank(s); every step ends with the all-gather of all index maps and C",
            "l2_policy": "inputs larger than L2 (%.1f MB of input per pair, >= 128 pairs per launch)" % per_pair_mb}


# ----------------------------------------------------------------------------------------------- data
def make_host_batch(pairs, seed=2000, pool=8, K=K_EIG):
    """Synthetic pairs: random unit features (the BASELINE feature model) and synthetic A-orthonormal bases
    (SURVEY.md 8d cfg2).  A small pool of distinct meshes is generated and pairs are drawn from it; every pair
    still owns its rows in the packed buffers, so memory traffic is that of distinct pairs."""
    from densematcher_b200 import synth
    from densematcher_b200.pipeline import PairBatchHost
    rng = np.random.default_rng(seed)
    bases = [synth.synthetic_basis(N_VERT, K, rng) for _ in range(pool)]
    feats = [synth.random_unit_features(N_VERT, D_FEAT, rng) for _ in range(pool)]
    ia, ib = rng.integers(0, pool, size=pairs), rng.integers(0, pool, size=pairs)
    ib = np.where(ib == ia, (ib + 1) % pool, ib)
    cat = lambda idx, f: np.concatenate([f(i) for i in idx])
    off = np.arange(pairs + 1, dtype=np.int64) * N_VERT
    b = PairBatchHost(
        F1=cat(ia, lambda i: feats[i]), F2=cat(ib, lambda i: feats[i]), off1=off, off2=off.copy(),
        Phi1=cat(ia, lambda i: bases[i][1]), Phi2=cat(ib, lambda i: bases[i][1]),
        evals1=np.stack([bases[i][0] for i in ia]), evals2=np.stack([bases[i][0] for i in ib]),
        area1=cat(ia, lambda i: bases[i][2]), area2=cat(ib, lambda i: bases[i][2]))
    b.pool = dict(bases=bases, feats=feats, ia=ia, ib=ib)
    return b


def make_host_bank(n_meshes, seed, K, sizes=None):
    """A pool of synthetic meshes as a MeshBankHost (cfg4: 8 meshes with a K = 200 basis)."""
    from densematcher_b200 import synth
    from densematcher_b200.pipeline import MeshBankHost
    rng = np.random.default_rng(seed)
    sizes = np.full(n_meshes, N_VERT) if sizes is None else np.asarray(sizes)
    bases = [synth.synthetic_basis(int(n), K, rng) for n in sizes]
    F = np.concatenate([synth.random_unit_features(int(n), D_FEAT, rng) for n in sizes])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return MeshBankHost(F=F, off=off, Phi=np.concatenate([b[1] for b in bases]), evals=np.stack([b[0] for b in bases]),
                        area=np.concatenate([b[2] for b in bases]))


def make_device_bank_cfg5(dev, seed=5000):
    """599 meshes in 24 categories generated ON THE DEVICE (a host QR per mesh would take minutes); returns
    (MeshBankDevice, category labels)."""
    import torch
    from densematcher_b200 import pipeline
    rng = np.random.default_rng(seed)
    n_meshes, n_cat, d, K = 599, 24, D_FEAT, K_EIG
    cats = np.sort(rng.integers(0, n_cat, size=n_meshes))
    sizes = rng.integers(1800, 2201, size=n_meshes)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    g = torch.Generator(device=dev).manual_seed(seed)
    F = torch.nn.functional.normalize(torch.randn(int(off[-1]), d, device=dev, generator=g), dim=1)
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
    return pipeline.MeshBankDevice(F, off, Phi=Phi, evals=evals, area=area, device=dev), cats


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_begin = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def wait_ready(self, timeout=3.0):
        """blocks until the first sample has arrived (NVML is up), at most ``timeout`` seconds"""
        t = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)
        return self

    def begin(self):
        """marks the start of the timed region: the process is started earlier (before the warm-up steps -- NVML takes
        100-300 ms to come up on a multi-GPU box, longer than a short timed region) and only the samples that arrive from
        here on are reported"""
        self.t_begin = time.perf_counter()
        return self

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.15)
        self.proc.terminate()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        inside = [r[1:] for r in self.rows if t0 <= r[0] <= t_end + 0.03]  # a sample is printed up to one period late
        self.rows = inside if inside else [r[1:] for r in self.rows if r[0] >= t0][:3]
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_pair_kdtree(F1, F2, P1, P2, a1, a2, ev1, ev2, n_jobs):
    """The reference path on the host for one pair, as the reference computes it: kd-tree NN both directions
    (knn_query, nn_utils.py:4-38), projection, closed-form C, FM_to_p2p with two kd-trees + the dense argmax."""
    from oracle import dm_oracle as orc
    orc.knn_query(F1, F2, n_jobs=n_jobs)
    orc.knn_query(F2, F1, n_jobs=n_jobs)
    A, B = orc.project(P1, a1, F1), orc.project(P2, a2, F2)
    C = orc.fmap_solve_closed_form(A, B, ev1, ev2, orc.fmap_c00(P1, P2, a1, a2), W_DESCR, W_LAP)
    emb2, emb1 = P2 @ C, P1 @ C.T
    orc.knn_query(emb2, P1, n_jobs=n_jobs)
    orc.knn_query(emb1, P2, n_jobs=n_jobs)
    orc.dense_argmax_override((emb2 @ P1.T) * a1[None, :])
    return C


def cpu_pair_bruteforce(F1, F2, P1, P2, a1, a2, ev1, ev2):
    """The strongest honest CPU baseline: the same outputs from float64 numpy GEMM + argmax
    (identical indices to the kd-tree) and the float64 closed form."""
    from oracle import dm_oracle as orc
    S = F2.astype(np.float64) @ F1.astype(np.float64).T
    S.argmax(1), S.argmax(0)
    A, B = orc.project(P1, a1, F1), orc.project(P2, a2, F2)
    C = orc.fmap_solve_closed_form(A, B, ev1, ev2, orc.fmap_c00(P1, P2, a1, a2), W_DESCR, W_LAP)
    emb2, emb1 = P2 @ C, P1 @ C.T
    S = emb2 @ P1.T
    (S - 0.5 * (emb1 * emb1).sum(1)[None, :]).argmax(1)
    (S - 0.5 * (emb2 * emb2).sum(1)[:, None]).argmax(0)
    (S * a1[None, :]).argmax(1), S.argmax(0)
    return C


def _pair_arrays(batch, p, k=K_EIG):
    s1, s2 = slice(batch.off1[p], batch.off1[p + 1]), slice(batch.off2[p], batch.off2[p + 1])
    return (batch.F1[s1], batch.F2[s2], batch.Phi1[s1][:, :k], batch.Phi2[s2][:, :k], batch.area1[s1], batch.area2[s2],
            batch.evals1[p][:k], batch.evals2[p][:k])


def cpu_baselines(cfg, batch, budget_s=12.0):
    """Both CPU legs on a bounded sample of the same workload; returns the `cpu_baseline` object."""
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    n_kd = 0
    while n_kd < min(8, batch.n_pairs) and (n_kd < 1 or time.perf_counter() - t0 < budget_s):
        cpu_pair_kdtree(*_pair_arrays(batch, n_kd), n_jobs=-1)
        n_kd += 1
    t_kd = (time.perf_counter() - t0) / n_kd
    cpu_pair_bruteforce(*_pair_arrays(batch, 0))  # warm the BLAS threads
    t0 = time.perf_counter()
    n_bf = 0
    while n_bf < min(8, batch.n_pairs) and (n_bf < 2 or time.perf_counter() - t0 < budget_s / 2):
        cpu_pair_bruteforce(*_pair_arrays(batch, n_bf))
        n_bf += 1
    t_bf = (time.perf_counter() - t0) / n_bf
    return {"value": 1.0 / t_kd, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_kd} pair(s) of the same batch through the oracle port of the reference path (the reference's "
                      "sklearn kd-tree call with n_jobs=-1, numpy/scipy float64)",
            "numpy_bruteforce": {"value": 1.0 / t_bf, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{n_bf} pairs, float64 numpy GEMM + argmax for every index map and the float64 "
                                           "closed form (BASELINE.md section 3: the strongest honest CPU baseline; same "
                                           "indices as the kd-tree)"}}


def run_reference_arm(args):
    """--impl reference: the CPU restatement of the reference path on this box's host cores, same metric / config.
    One step = `s` pair(s) of the workload (distinct pairs from step to step), `s` chosen from the first measured pair so
    that the K + W steps end within about two minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg = args.config
    pairs_cfg = args.pairs or CONFIG_DEFAULTS[cfg]["pairs"] or 14964
    if cfg in ("cfg2a", "cfg2b", "cfg5", "cfg4"):
        batch = make_host_batch(8, pool=4, K=200 if cfg in ("cfg4", "cfg2b") else K_EIG)
        if cfg in ("cfg2a", "cfg5"):
            one = lambda p: cpu_pair_kdtree(*_pair_arrays(batch, p % 8), n_jobs=-1)
            what = "oracle port of the reference path: sklearn kd-tree n_jobs=-1 + numpy/scipy float64"
        else:
            from oracle import dm_oracle as orc
            k0, k1 = (30, 100) if cfg == "cfg2b" else (30, 200)

            def one(p):
                F1, F2, P1, P2, a1, a2, ev1, ev2 = _pair_arrays(batch, p % 8, k=200)
                C = (cpu_pair_kdtree(F1, F2, P1[:, :K_EIG], P2[:, :K_EIG], a1, a2, ev1[:K_EIG], ev2[:K_EIG], -1)[:k0, :k0]
                     if cfg == "cfg2b" else np.eye(k0))
                orc.zoomout_refine(C, P1, P2, nit=k1 - k0, step=1, A2=a2, return_p2p=True)
            what = ("oracle port with the float64 BRUTE-FORCE nearest neighbour inside the ZoomOut ladder (identical results; "
                    "the reference's kd-tree ladder takes ~137 s per pair, SURVEY section 6)")
    else:  # cfg3: NN only, ragged
        from oracle import dm_oracle as orc
        from densematcher_b200 import synth
        rng = np.random.default_rng(3000)
        sizes = rng.integers(1500, 2501, size=(8, 2))
        feats = [(synth.random_unit_features(int(a), D_FEAT, rng), synth.random_unit_features(int(b), D_FEAT, rng)) for a, b in sizes]

        def one(p):
            F1, F2 = feats[p % 8]
            orc.knn_query(F1, F2, n_jobs=-1)
            orc.knn_query(F2, F1, n_jobs=-1)
        what = "the reference's kd-tree knn_query, both directions, n_jobs=-1"
    t0 = time.perf_counter()
    one(0)
    t_pair = time.perf_counter() - t0
    n_steps = max(1, args.steps) + max(0, args.warmup)
    s = int(max(1, min(8, 120.0 / (n_steps * t_pair))))
    p = 1
    for _ in range(max(0, args.warmup - 1)):
        for _ in range(s):
            one(p); p += 1
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        for _ in range(s):
            one(p); p += 1
    dt = time.perf_counter() - t0
    val = s * max(1, args.steps) / dt
    desc = f"{s} pair(s) of the workload per step, {max(1, args.steps)} steps over distinct pairs ({what})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": CONFIG_DEFAULTS[cfg]["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, pairs_cfg, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore its pinned-memory allocations, first-touch) to the CPUs of the NUMA node its
    GPU hangs off: with one rank per GPU the host->device copies of 8 ranks otherwise cross the socket interconnect.
    Best effort: returns a description, or a reason string starting with "unbound"."""
    try:
        import torch
        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                           capture_output=True, text=True, timeout=10).stdout.strip().lower()
        # nvidia-smi prints an 8-digit domain ("00000000:1b:00.0"), sysfs uses four
        dom, rest = q.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
            if n_nodes <= 1:
                return "unbound: single NUMA node"
            node = local_rank * n_nodes // max(1, torch.cuda.device_count())  # no affinity exposed: spread the ranks
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"unbound: no allowed cpu on node {node}"
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}, {len(allowed)} cpus"
    except Exception as e:  # noqa: BLE001
        return f"unbound: {type(e).__name__}: {e}"[:160]


# ----------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--config", default="cfg2a", choices=sorted(CONFIG_DEFAULTS))
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU per step (weak configs) / in total (cfg4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--verify", type=int, default=0, help="cfg4: check this many pairs per rank against the CPU oracle after the run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    heavy = args.config in ("cfg4", "cfg5", "cfg2b")
    if args.steps is None:
        args.steps = 1 if args.config == "cfg4" else 3 if heavy else 10
    if args.warmup is None:
        args.warmup = 1 if args.config == "cfg4" else 3
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.config != "cfg4":
        args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from densematcher_b200 import _lib, nn as dnn, fm as dfm, pipeline

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    cfg = args.config
    P = args.pairs or CONFIG_DEFAULTS[cfg]["pairs"]
    i32 = torch.int32

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    comm_stream = torch.cuda.Stream(device) if world > 1 else None
    gather_state = {"pad": None, "bufs": [None, None], "i": 0, "peer": None, "mode": None}
    gather_mode = os.environ.get("DM_BENCH_GATHER", "peer")   # peer (NVLink copy engines) | nccl | none (diagnostics only)

    def gather_all(res):
        """The path's one collective, EVERY step: all index maps and C of all shards.  The results of a step are packed
        into one int32 buffer (C viewed as int32 words; ragged shards padded to the largest shard, whose size is agreed
        once) and gathered on a side stream, so that the NVLink transfer of step i overlaps the kernels of step i + 1 (two
        buffer sets alternate; the timed region ends after the last gather).  Default transport: pipeline.PeerGather
        (peer-memory writes by the copy engines + a device-side barrier); NCCL's all_gather_into_tensor when symmetric
        memory is unavailable or DM_BENCH_GATHER=nccl."""
        if world == 1 or gather_mode == "none":
            return
        parts = [t.reshape(-1).view(torch.int32) for n_, t in sorted(res.items()) if n_ != "status" and t is not None]
        total = sum(p_.numel() for p_ in parts)
        if gather_state["pad"] is None:
            n = torch.tensor([total], device=device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX)
            gather_state["pad"] = int(n.item())
            gather_state["mode"] = "nccl"
            if gather_mode == "peer":
                ok = torch.ones(1, device=device)
                try:
                    gather_state["peer"] = pipeline.PeerGather(gather_state["pad"], device)
                except Exception as e:  # noqa: BLE001 -- any failure of the symmetric-memory setup: every rank falls back
                    ok.zero_()
                    if rank == 0:
                        print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); NCCL all-gather", file=sys.stderr)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok.item()) > 0:
                    gather_state["mode"] = "peer"
                else:
                    gather_state["peer"] = None
        m = gather_state["pad"]
        i = gather_state["i"] = gather_state["i"] ^ 1
        if gather_state["bufs"][i] is None:
            gather_state["bufs"][i] = (torch.zeros(m, dtype=torch.int32, device=device),
                                       None if gather_state["mode"] == "peer" else
                                       torch.empty(world * m, dtype=torch.int32, device=device), torch.cuda.Event())
        send, recv, done = gather_state["bufs"][i]
        cur = torch.cuda.current_stream(device)
        cur.wait_event(done)                       # the gather that last used this buffer set has finished
        torch.cat(parts, out=send[:total])
        if gather_state["mode"] == "peer":
            pg = gather_state["peer"]
            pg.gather(send)
            done.record(pg.stream)
            return
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ready)
            dist.all_gather_into_tensor(recv, send)
            done.record(comm_stream)

    def finish_gathers():
        if world > 1:
            torch.cuda.current_stream(device).wait_stream(comm_stream)
            if gather_state["peer"] is not None:
                gather_state["peer"].wait()

    host = dev = bank = None
    extra = {}
    launches_per_step = None
    # ------------------------------------------------------------------ workloads
    if cfg in ("cfg2a", "cfg2b"):
        host = make_host_batch(P, seed=2000 + rank, K=K_EIG).pin()
        dev = host.to_device(device)
        if cfg == "cfg2a":
            def step():
                res = pipeline.match_pairs_device(dev, k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP, check=False)
                gather_all(res)
                return res
            # launches of OUR kernels per step (profiles/launches_r2_end_step.csv): feature NN 8 (2 prep, 2 per-pair maxima,
            # score, column finalise, 2 re-evaluation) + projection 2 x 3 + pinned entry 1 + solve 6 (2 Gram GEMMs, float32
            # pack, factor/refine, lazy float64 pack, float64 fallback) + FM->p2p 16 (2 eigenbasis splits enqueued beside the
            # solve, per-row arrays of the database side + per-pair maxima, split of C, 2 embeddings, score pass, column
            # finalise, on-demand fill, 2 flagged GEMMs + 2 bias passes, 2 re-evaluation)
            launches_per_step = 8 + 6 + 1 + 6 + 16
        else:
            nit = 70

            def step():
                res = pipeline.match_pairs_device(dev, k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP, check=False)
                C0 = res["C"][:, :30, :30].contiguous()
                Cz, pz = dfm.zoomout(C0, dev.Phi1, dev.Phi2, dev.area2, nit, 1, dev.o1, dev.o2, return_p2p=True, out_dtype=i32)
                res.update(C_zo=Cz, p2p_zo=pz)
                gather_all(res)
                return res
            launches_per_step = 36 + nit * 11 + 11     # 11 launches per rung (profiles/r2_zoomout_launches.md minus the merged bookkeeping kernels)
        units_per_rank = P
    elif cfg == "cfg3":
        g = torch.Generator(device=device).manual_seed(3000 + rank)
        rng = np.random.default_rng(3000 + rank)
        nq, nd = rng.integers(1500, 2501, size=P), rng.integers(1500, 2501, size=P)
        qo, do = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64), np.concatenate([[0], np.cumsum(nd)]).astype(np.int64)
        Y = torch.nn.functional.normalize(torch.randn(int(qo[-1]), D_FEAT, device=device, generator=g), dim=1)
        X = torch.nn.functional.normalize(torch.randn(int(do[-1]), D_FEAT, device=device, generator=g), dim=1)
        qoff, doff = dnn.Offsets(torch.from_numpy(qo).to(device), qo), dnn.Offsets(torch.from_numpy(do).to(device), do)

        def step():
            (r,), (c,) = dnn.nn_argmax(Y, X, qoff, doff, row_epi=(dnn.COSINE_UNIT,), col_epi=(dnn.COSINE_UNIT,), out_dtype=i32)
            res = dict(nn_p2p_21=r, nn_p2p_12=c)
            gather_all(res)
            return res
        launches_per_step = 8
        units_per_rank = P
        extra["nn_algorithmic_gflop_per_step_per_gpu"] = 2.0 * float(np.sum(nq.astype(np.float64) * nd)) * D_FEAT / 1e9
    elif cfg == "cfg4":
        nit, k0, K = 170, 30, 200
        hbank = make_host_bank(8, seed=4000 + rank, K=K).pin()
        bank = pipeline.MeshBankDevice(hbank.F, hbank.off, Phi=hbank.Phi, evals=hbank.evals, area=hbank.area, device=device)
        lo, hi = pipeline.shard_pairs(P, rank, world)
        rng = np.random.default_rng(4100)
        src_all, dst_all = rng.integers(0, 8, size=P), rng.integers(0, 8, size=P)
        dst_all = np.where(dst_all == src_all, (dst_all + 1) % 8, dst_all)
        chunk = 128
        zeros_ev = torch.zeros(chunk, k0, dtype=torch.float64, device=device)

        def run_chunk(a, b):
            pb = bank.assemble(src_all[a:b], dst_all[a:b])
            n = b - a
            A = dfm.project(pb.Phi1, pb.area1, pb.F1, pb.o1, k=k0)
            B = dfm.project(pb.Phi2, pb.area2, pb.F2, pb.o2, k=k0)
            # C0 = B A^+ (north star): the closed form with w_lap = 0; its pinned first column as in the reference's x0
            C0 = dfm.fmap_solve(A, B, zeros_ev[:n] + torch.arange(k0, device=device), zeros_ev[:n] + torch.arange(k0, device=device),
                                pipeline.fmap_c00(pb), 1.0, 0.0, check=False)
            Cz, pz = dfm.zoomout(C0, pb.Phi1, pb.Phi2, pb.area2, nit, 1, pb.o1, pb.o2, return_p2p=True, out_dtype=i32)
            return pb, C0, Cz, pz

        def step():
            Cs, ps = [], []
            for a in range(lo, hi, chunk):
                _, _, Cz, pz = run_chunk(a, min(hi, a + chunk))
                Cs.append(Cz); ps.append(pz)
            res = dict(C_zo=torch.cat(Cs), p2p_zo=torch.cat(ps))
            gather_all(res)
            return res
        n_chunks = (hi - lo + chunk - 1) // chunk
        launches_per_step = n_chunks * (12 + nit * 11 + 11)
        units_per_rank = hi - lo
    else:  # cfg5
        bank, cats = make_device_bank_cfg5(device)
        src, dst = pipeline.intra_category_pairs(cats)
        P = len(src)
        lo, hi = pipeline.shard_pairs(P, rank, world)
        kw = dict(k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP, out_dtype=i32, check=False)

        def step():
            # a step is the WHOLE job: the once-per-mesh preparation of the bank (operand splits, norms, projections:
            # dm_bank_prepare) is redone inside every step, then every pair of this rank goes through dm_match_bank_pairs
            for st_ in bank._states.values():
                bank._state_buf = st_.state
            bank._states.clear()
            outs, _ = pipeline.match_bank_pairs(bank, src, dst, chunk_pairs=128, rank=rank, world=world, to_host=False, **kw)
            res = {n: torch.cat([o[n] for o in outs]) for n in outs[0] if n != "status"}
            gather_all(res)
            return res
        launches_per_step = ((hi - lo + 127) // 128) * 36 + 2 + 3 * ((bank.n_meshes + 127) // 128)
        units_per_rank = hi - lo
    torch.cuda.synchronize()

    # ------------------------------------------------------------------ timed steps (device-resident inputs)
    sampler = ClockSampler(local_rank).start().wait_ready() if rank == 0 else None   # streaming before the warm-up, see begin()
    for _ in range(args.warmup):
        res = step()
    barrier()
    if sampler:
        sampler.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    finish_gathers()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([ms], device=device)
    units = torch.tensor([float(units_per_rank)], device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    ms = float(tmax.item())
    total_units = float(units.item())
    value = total_units * args.steps / (ms * 1e-3)
    if "status" in res and res["status"] is not None and int(res["status"][0]) != 0:
        raise SystemExit("bench: a functional-map system was singular")

    # ------------------------------------------------------------------ rooflines: the kernels that dominate the step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    which = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    info = lib.dm_build_info().decode()

    def event_ms(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    roof, top = None, []
    reps = max(3, min(args.steps, 10))
    if cfg in ("cfg2a", "cfg2b", "cfg3"):
        skip = _lib.DM_SKIP_PREP | _lib.DM_SKIP_FINISH
        if cfg == "cfg3":
            nn_call = lambda fl: dnn.nn_argmax(Y, X, qoff, doff, row_epi=(dnn.COSINE_UNIT,), col_epi=(dnn.COSINE_UNIT,), flags=fl, out_dtype=i32)
            n_pairs_k, alg_flops = P, extra["nn_algorithmic_gflop_per_step_per_gpu"] * 1e9
            alg_bytes = float((qo[-1] + do[-1]) * (4 * D_FEAT + 4))
        else:
            nn_call = lambda fl: dnn.nn_argmax(dev.F2, dev.F1, dev.off2, dev.off1, row_epi=(dnn.COSINE_UNIT,),
                                               col_epi=(dnn.COSINE_UNIT,), max_q=dev.max2, max_db=dev.max1, flags=fl, out_dtype=i32)
            n_pairs_k, alg_flops, alg_bytes = P, P * ALG_FLOPS_NN, P * ALG_BYTES_NN
        nn_call(0)
        kern_ms = event_ms(lambda: nn_call(skip), reps)
        nn_stage_ms = event_ms(lambda: nn_call(0), reps)
        alg_tflops = alg_flops / (kern_ms * 1e-3) / 1e12
        hbm_gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
        ex = 3 * alg_tflops
        roof = {"bound": "tensor", "achieved": ex, "peak": tf_peak, "unit": "TFLOP/s", "frac": ex / tf_peak,
                "traffic": NCU_DRAM_BYTES_PER_PAIR * n_pairs_k if cfg != "cfg3" else None,
                "kernel": "nn_tc_kernel<1,1> feature-NN score pass (tcgen05, CTA pair)", "kernel_ms": kern_ms,
                "nn_stage_ms": nn_stage_ms, "algorithmic_tflops": alg_tflops, "hbm_gbs": hbm_gbs, "hbm_peak_gbs": hbm_peak,
                "hbm_frac": hbm_gbs / hbm_peak, "peaks": which,
                "note": "3 bf16 MMA passes per fp32-grade product (split-bf16); the pass is tensor-bound (AI ~ 500 flop/B, "
                        "SURVEY.md 8d), hbm_frac is the figure BASELINE.json asks for; traffic = dram read+write bytes per "
                        "launch from the ncu --set full capture profiles/r2_nn_tc_full_raw.csv (6.82 MB per pair vs 6.16 "
                        "MB algorithmic)"}
        top.append(dict(roof))
    if cfg in ("cfg2a", "cfg2b"):
        k = K_EIG
        A = dfm.project(dev.Phi1, dev.area1, dev.F1, dev.o1, k=k)
        B = dfm.project(dev.Phi2, dev.area2, dev.F2, dev.o2, k=k)
        c00 = pipeline.fmap_c00(dev)
        solve = lambda: dfm.fmap_solve(A, B, dev.evals1[:, :k], dev.evals2[:, :k], c00, W_DESCR, W_LAP, check=False)
        C = solve()
        solve_stage_ms = event_ms(solve, reps)
        os.environ["DM_SOLVE_SKIP_PREP"] = "1"
        solve_kern_ms = event_ms(solve, reps)
        os.environ.pop("DM_SOLVE_SKIP_PREP")
        n_sys, n = P * k, k - 1
        # factorisation n^3/3 + one refinement step: residual 2 n^2 (float64) + three triangular solve pairs ~ 6 n^2
        flops = n_sys * (n ** 3 / 3.0 + 8.0 * n * n)
        ach = flops / (solve_kern_ms * 1e-3) / 1e12
        top.append({"kernel": "fmap_solve32w_kernel (float32 Cholesky + float64 refinement, one warp per system)",
                    "kernel_ms": solve_kern_ms, "stage_ms": solve_stage_ms, "bound": "fp32_fma", "achieved": ach,
                    "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS,
                    "traffic": NCU_SOLVE_DRAM_BYTES_PER_PAIR * P,
                    "note": f"{n_sys} SPD systems of size {n}; algorithmic flops n^3/3 + 8 n^2 each; peak = 148 SM x 128 FFMA/clk "
                            "x 1.965 GHz; the kernel is latency-bound (9-10 resident warps per SM: 22 KB of shared memory per "
                            "system), see DESIGN.md 5.3"})
        ALL = ("p2p_21", "p2p_12", "dense_21", "dense_12")
        f2p = lambda fl: dfm.fm_to_p2p(C, dev.Phi1[:, :k], dev.Phi2[:, :k], dev.area1, dev.o1, dev.o2, want=ALL, flags=fl, out_dtype=i32)
        f2p(0)
        f2p_kern_ms = event_ms(lambda: f2p(_lib.DM_SKIP_PREP | _lib.DM_SKIP_FINISH), reps)
        f2p_stage_ms = event_ms(lambda: f2p(0), reps)
        kp = (k + 63) // 64 * 64
        ex = 3 * P * 2.0 * N_VERT * N_VERT * kp / (f2p_kern_ms * 1e-3) / 1e12
        red = P * 4.0 * N_VERT * N_VERT / (f2p_kern_ms * 1e-3)           # score reductions per second (4 index maps)
        alu_peak = ALU_PIPE_OPS_PER_S / 4.0                              # 4 ALU-pipe instructions per tracked score (1 LOP3 key + 3 FMNMX top-2)
        top.append({"kernel": "f2p_tc_kernel<2> FM->p2p score pass (dual accumulator, 4 index maps from one pass)", "kernel_ms": f2p_kern_ms,
                    "stage_ms": f2p_stage_ms, "bound": "alu", "achieved": red / 1e12, "peak": alu_peak / 1e12,
                    "unit": "T score-reductions/s", "frac": red / alu_peak, "tensor_frac": ex / tf_peak,
                    "traffic": NCU_F2P_DRAM_BYTES_PER_PAIR * P,
                    "note": "ALU-pipe bound, not tensor bound (ncu profiles/r2_kernels.md: ALU pipe 70 %% busy, tensor 45 %%): 4 argmax reductions "
                            "over every score; peak model = half-rate ALU pipe / 4 instructions per tracked score (LOP3 key + 3 "
                            "FMNMX); tensor_frac = executed bf16 flops (3 "
                            "passes, K padded to %d) / measured bf16 peak" % kp})
        top.sort(key=lambda r: -r["kernel_ms"])

    # ------------------------------------------------------------------ end to end through the host-buffer entries
    e2e = None
    e2e_bank = None
    if not args.no_e2e and cfg in ("cfg2a", "cfg3"):
        if cfg == "cfg2a":
            kw = dict(k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP, chunk_pairs=max(8, P // 8), copy=False)
            hb = host
        else:
            hb = pipeline.PairBatchHost(F1=X.cpu().numpy(), F2=Y.cpu().numpy(), off1=do, off2=qo).pin()
            kw = dict(functional_map=False, chunk_pairs=128, copy=False)
        e2e_call = lambda: pipeline.match_pairs_host(hb, device, **kw)
        for _ in range(2):  # both alternating sets of pinned result buffers exist before the timed region
            out = e2e_call()
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            out = e2e_call()
        g1.record()
        barrier()
        e2e_ms = max(g0.elapsed_time(g1), 1e3 * (time.perf_counter() - t0))
        t2 = torch.tensor([e2e_ms], device=device)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        d2h = int(sum(v.nbytes for v in out.values()))
        e2e = {"value": world * P * n_e2e / (float(t2.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": hb.h2d_bytes(), "d2h_bytes_per_step": d2h, "steps": n_e2e,
               "entry": "pipeline.match_pairs_host: every pair's features / basis / areas cross PCIe (pair-shaped input)"}
    if not args.no_e2e and cfg == "cfg2a":
        # bank-shaped entry on the SAME pairs: the batch draws its pairs from a pool of 8 meshes; a dataset-shaped caller
        # (cfg5: every mesh takes part in ~24 pairs) uploads each mesh once and names the pairs by id
        from densematcher_b200.pipeline import MeshBankHost
        pool = host.pool
        hbank = MeshBankHost(F=np.concatenate(pool["feats"]), off=np.arange(9, dtype=np.int64) * N_VERT,
                             Phi=np.concatenate([b[1] for b in pool["bases"]]), evals=np.stack([b[0] for b in pool["bases"]]),
                             area=np.concatenate([b[2] for b in pool["bases"]])).pin()
        kwb = dict(k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP, chunk_pairs=P, copy=False)   # one launch set per call (19.3 k vs 14.6 k pairs/s at P // 4)
        bank_call = lambda: pipeline.match_bank_pairs_host(hbank, pool["ia"], pool["ib"], device, **kwb)
        for _ in range(2):
            outb = bank_call()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            outb = bank_call()
        torch.cuda.synchronize()
        tb = torch.tensor([1e3 * (time.perf_counter() - t0)], device=device)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        e2e_bank = {"value": world * P * n_e2e / (float(tb.item()) * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": hbank.h2d_bytes() + 16 * P,
                    "d2h_bytes_per_step": int(sum(v.nbytes for n_, v in outb.items() if not n_.startswith("off"))),
                    "steps": n_e2e,
                    "entry": "pipeline.match_bank_pairs_host: the 8 distinct meshes of the batch cross PCIe once per step, the "
                             "128 pairs are id lists; operand splits / norms / projections are made once per mesh (dm_bank_prepare) and the "
                             "pair kernels read the meshes' rows in place (dm_match_bank_pairs): dataset-shaped input, cfg5"}
    if not args.no_e2e and cfg in ("cfg4", "cfg5") and world == 1 and cfg == "cfg4":
        pass  # cfg4 / cfg5 e2e: the bank is uploaded once for the whole job (0.1 % of the step): reported by cfg2a's e2e_bank

    # ------------------------------------------------------------------ cfg4: oracle check of a seeded sample
    verify = None
    if cfg == "cfg4" and args.verify > 0:
        from oracle import dm_oracle as orc  # the checker, outside every timed region
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=max(1, (os.cpu_count() or 8) // world))  # the ranks check their samples side by side
        nv = min(args.verify, units_per_rank)
        pb, C0, Cz, pz = run_chunk(lo, lo + nv)
        torch.cuda.synchronize()
        worst, exact = 0.0, True
        for i in range(nv):
            s1, s2 = slice(pb.off1_h[i], pb.off1_h[i + 1]), slice(pb.off2_h[i], pb.off2_h[i + 1])
            Co, po = orc.zoomout_refine(C0[i].cpu().numpy(), pb.Phi1[s1].cpu().numpy(), pb.Phi2[s2].cpu().numpy(), nit=nit,
                                        step=1, A2=pb.area2[s2].cpu().numpy(), return_p2p=True)
            worst = max(worst, float(np.linalg.norm(Cz[i].cpu().numpy() - Co) / np.linalg.norm(Co)))
            exact = exact and bool(np.array_equal(pz[s2].cpu().numpy(), po))
        v = torch.tensor([worst, 0.0 if exact else 1.0], device=device)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        verify = {"pairs_per_rank": nv, "max_relF_C": float(v[0].item()), "final_p2p_exact": bool(v[1].item() == 0.0),
                  "bar": "C <= 1e-4 relative Frobenius, final p2p identical to the float64 oracle ladder"}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and cfg == "cfg2a":
        cpu = cpu_baselines(cfg, host)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": CONFIG_DEFAULTS[cfg]["scaling"], "vs_baseline": None,
                "dtype": "f32 scores + f64 re-evaluation / f64 functional map", "data": "synthetic",
                "config": workload_config(cfg, P, world), "roofline": roof, "roofline_top": top, "cpu_baseline": cpu,
                "e2e": e2e, "e2e_bank": e2e_bank, "gpu_launches": (launches_per_step or 0) * args.steps, "clocks": clocks,
                "engine": "tcgen05", "lib": info, "numa_binding": numa, "gather_transport": gather_state["mode"] if world > 1 else None, "verify": verify, **extra}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
