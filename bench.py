#!/usr/bin/env python
"""Headline benchmark: mesh pairs matched per second (N = M = 2000 vertices, d = 384 features, k = 100 LBO basis).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl ours|reference]

One "step" = one pass of the correspondence hot path over a batch of B synthetic mesh pairs per GPU
(BASELINE.json configs[1] replicated B times; inputs 9.4 MB/pair, so the batch is far larger than L2):
  feature NN (cosine argmax, both directions, one fused pass) -> projection Phi^T A F (both meshes) ->
  closed-form C (k = 100) -> FM->p2p (kd-tree-equivalent pair + dense-argmax pair from one pass).
Prints ONE JSON line (rank 0).  `value` = pairs/s with inputs resident in HBM; `e2e` = the same through the
host-buffer entry (pinned H2D of every input + D2H of every result inside the timed region).
`--impl reference` times the CPU restatement of the reference path (oracle/, sklearn kd-tree like the
reference) on a bounded sample.  Multi-GPU: pairs shard across ranks (weak scaling), one final all-gather.
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
NCU_DRAM_BYTES_PER_PAIR = (400.146432e6 + 38.645248e6) / 64      # ncu --set full capture of nn_tc_kernel<1,1,0,1> (CTA-pair mode) at 64 pairs


def workload_config(pairs, n_gpus):
    return {"workload": f"cfg2a x{pairs}/GPU: pairs of N=M={N_VERT} meshes, d={D_FEAT} unit features, k={K_EIG} LBO basis; "
                        "feature NN (both directions) + projection + closed-form C + FM->p2p (4 index maps)",
            "pairs_per_gpu": pairs, "n": N_VERT, "d": D_FEAT, "k": K_EIG, "w_descr": W_DESCR, "w_lap": W_LAP,
            "parallelism": f"pairs sharded over {n_gpus} rank(s), final all-gather of the index maps",
            "l2_policy": "inputs larger than L2 (%.0f MB per step per GPU)" % (pairs * 9.4)}


# ----------------------------------------------------------------------------------------------- data
def make_host_batch(pairs, seed=2000, pool=8):
    """Synthetic pairs: random unit features (the BASELINE feature model) and synthetic A-orthonormal bases
    (SURVEY.md 8d cfg2).  A small pool of distinct meshes is generated and pairs are drawn from it; every pair
    still owns its rows in the packed buffers, so memory traffic is that of distinct pairs."""
    from oracle import meshgen
    from densematcher_b200.pipeline import PairBatchHost
    rng = np.random.default_rng(seed)
    bases = [meshgen.synthetic_basis(N_VERT, K_EIG, rng) for _ in range(pool)]
    feats = [meshgen.random_unit_features(N_VERT, D_FEAT, rng) for _ in range(pool)]
    ia, ib = rng.integers(0, pool, size=pairs), rng.integers(0, pool, size=pairs)
    ib = np.where(ib == ia, (ib + 1) % pool, ib)
    cat = lambda idx, f: np.concatenate([f(i) for i in idx])
    off = np.arange(pairs + 1, dtype=np.int64) * N_VERT
    return PairBatchHost(
        F1=cat(ia, lambda i: feats[i]), F2=cat(ib, lambda i: feats[i]), off1=off, off2=off.copy(),
        Phi1=cat(ia, lambda i: bases[i][1]), Phi2=cat(ib, lambda i: bases[i][1]),
        evals1=np.stack([bases[i][0] for i in ia]), evals2=np.stack([bases[i][0] for i in ib]),
        area1=cat(ia, lambda i: bases[i][2]), area2=cat(ib, lambda i: bases[i][2]))


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_pair_time(batch, n_pairs, n_jobs):
    """The reference path on the host for `n_pairs` pairs of the batch: kd-tree NN both directions
    (knn_query, nn_utils.py:4-38), projection, closed-form C, FM_to_p2p with kd-trees + dense argmax."""
    from oracle import dm_oracle as orc
    t0 = time.perf_counter()
    for p in range(n_pairs):
        s1, s2 = slice(batch.off1[p], batch.off1[p + 1]), slice(batch.off2[p], batch.off2[p + 1])
        F1, F2 = batch.F1[s1], batch.F2[s2]
        orc.knn_query(F1, F2, n_jobs=n_jobs)
        orc.knn_query(F2, F1, n_jobs=n_jobs)
        P1, P2, a1, a2 = batch.Phi1[s1], batch.Phi2[s2], batch.area1[s1], batch.area2[s2]
        A, B = orc.project(P1, a1, F1), orc.project(P2, a2, F2)
        C = orc.fmap_solve_closed_form(A, B, batch.evals1[p], batch.evals2[p], orc.fmap_c00(P1, P2, a1, a2), W_DESCR, W_LAP)
        k2, k1 = C.shape
        emb2, emb1 = P2 @ C, P1 @ C.T
        orc.knn_query(emb2, P1, n_jobs=n_jobs)
        orc.knn_query(emb1, P2, n_jobs=n_jobs)
        orc.dense_argmax_override((emb2 @ P1.T) * a1[None, :])
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 2
    batch = make_host_batch(sample, pool=4)
    pairs_cfg = args.pairs
    for _ in range(min(args.warmup, 1)):
        cpu_pair_time(batch, 1, -1)
    times = [cpu_pair_time(batch, sample, -1) for _ in range(max(1, min(args.steps, 3)))]
    dt = max(times) if len(times) < 3 else float(np.median(times))
    val = sample / dt
    desc = f"{sample} pairs of the workload per step (oracle port of the reference path: sklearn kd-tree n_jobs=-1 + numpy/scipy float64)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(pairs_cfg, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore its pinned-memory allocations, first-touch) to the CPUs of the NUMA node its
    GPU hangs off: with one rank per GPU the host->device copies of 8 ranks otherwise cross the socket interconnect.
    Best effort: returns a description or None when the topology cannot be read."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return f"numa node {node}, {len(allowed)} cpus"
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=128, help="pairs per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from densematcher_b200 import _lib, nn as dnn, pipeline

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    P = args.pairs
    host = make_host_batch(P, seed=2000 + rank).pin()
    dev = host.to_device(device)
    torch.cuda.synchronize()

    def step():
        return pipeline.match_pairs_device(dev, k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step()
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    if world > 1:  # the path's one collective: gather the index maps of all shards
        counts = [res["p2p_21"].numel()] * world
        pipeline.gather_results(res["p2p_21"], counts)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * P * args.steps / (ms * 1e-3)

    # ---- dominant kernel alone: the fused feature-NN score pass (one launch per call with both phase-skip flags)
    skip = _lib.DM_SKIP_PREP | _lib.DM_SKIP_FINISH
    nn_call = lambda fl: dnn.nn_argmax(dev.F2, dev.F1, dev.off2, dev.off1, row_epi=(dnn.COSINE_UNIT,),
                                       col_epi=(dnn.COSINE_UNIT,), max_q=dev.max2, max_db=dev.max1, flags=fl,
                                       out_dtype=torch.int32)
    nn_call(0)
    for _ in range(2):
        nn_call(skip)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, args.steps)
    k0.record()
    for _ in range(reps):
        nn_call(skip)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / reps
    # whole NN stage (prep + score pass + column finalise + float64 re-evaluation)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps):
        nn_call(0)
    s1.record()
    torch.cuda.synchronize()
    nn_stage_ms = s0.elapsed_time(s1) / reps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    which = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    info = _lib.load().dm_build_info().decode()
    engine = "tcgen05" if "tcgen05" in info and not os.environ.get("DM_FORCE_FFMA") else "ffma"
    alg_tflops = P * ALG_FLOPS_NN / (kern_ms * 1e-3) / 1e12
    hbm_gbs = P * ALG_BYTES_NN / (kern_ms * 1e-3) / 1e9
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12   # CUDA-core FFMA peak at max clock, TFLOP/s
    if engine == "ffma":
        roof = {"bound": "fp32_ffma", "achieved": alg_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": alg_tflops / fp32_peak, "traffic": None,
                "note": "score pass is math-bound (AI ~ 500 flop/B, SURVEY.md 8d); peak = 148 SM x 128 FFMA/clk x 1.965 GHz"}
    else:
        ex = 3 * alg_tflops
        roof = {"bound": "tensor", "achieved": ex, "peak": tf_peak, "unit": "TFLOP/s", "frac": ex / tf_peak,
                "traffic": NCU_DRAM_BYTES_PER_PAIR * P,
                "note": "3 bf16 MMA passes per fp32-grade product (split-bf16); peak " + which +
                        "; traffic = dram read+write bytes of this kernel per launch from the ncu --set full capture "
                        "profiles/r1_end_nn_tc_full_raw.csv (6.86 MB per pair vs 6.16 MB algorithmic)"}
    roof.update({"kernel": "nn score pass (" + engine + ")", "kernel_ms": kern_ms, "nn_stage_ms": nn_stage_ms,
                 "algorithmic_tflops": alg_tflops, "hbm_gbs": hbm_gbs, "hbm_peak_gbs": hbm_peak,
                 "hbm_frac": hbm_gbs / hbm_peak, "peaks": which})

    # ---- end to end through the host-buffer entry
    e2e = None
    if not args.no_e2e:
        kw = dict(k=K_EIG, w_descr=W_DESCR, w_lap=W_LAP, chunk_pairs=max(8, P // 8), copy=False)
        for _ in range(2):  # both alternating sets of pinned result buffers exist before the timed region
            out = pipeline.match_pairs_host(host, device, **kw)
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            out = pipeline.match_pairs_host(host, device, **kw)
        g1.record()
        barrier()
        e2e_ms = max(g0.elapsed_time(g1), 1e3 * (time.perf_counter() - t0))
        t2 = torch.tensor([e2e_ms], device=device)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        d2h = int(sum(v.nbytes for v in out.values()))
        e2e = {"value": world * P * n_e2e / (float(t2.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": host.h2d_bytes(), "d2h_bytes_per_step": d2h, "steps": n_e2e}

    # ---- CPU baseline beside it (rank 0, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 2
        dt = cpu_pair_time(host, sample, -1)
        cpu = {"value": sample / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"{sample} pairs of the same batch, oracle port of the reference path "
                         "(sklearn kd-tree n_jobs=-1, numpy/scipy float64)"}

    # launches of OUR kernels per step (profiles/launches_*; one dm_match_pairs call): feature NN 8 (2 prep, 2 per-pair
    # maxima, score, column finalise, 2 re-evaluation) + projection 2 x 3 (split of Phi, tcgen05 GEMM, reduce; the
    # feature splits come from the NN stage) + pinned entry 1 + solve 4 (2 Gram GEMMs, pack, Cholesky) + FM->p2p 11
    # (2 embedding GEMMs, norms, 2 prep, 2 per-pair maxima, score, finalise, 2 re-evaluation)
    launches_per_step = 8 + 6 + 1 + 4 + 11
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 scores + f64 re-evaluation / f64 functional map", "data": "synthetic",
                "config": workload_config(P, world), "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "engine": engine, "lib": info, "numa_binding": numa}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
