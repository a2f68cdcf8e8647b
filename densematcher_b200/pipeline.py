"""The batched correspondence pipeline: what ``compute_surface_map`` does per pair
(densematcher/functional_map.py:44-50), for a ragged batch of pairs resident in HBM.

    feature NN   : p2p by cosine / Euclidean argmax of the feature similarity (both directions, one pass)
    projection   : A = Phi1^T A1 F1, B = Phi2^T A2 F2            (base_functions.py:526-532)
    solve        : closed-form C of the descriptor + Laplacian energy (functional.py:352-487)
    FM -> p2p    : kd-tree-equivalent p2p_21 / p2p_12 and the dense-argmax override (convert.py:96-147,
                   functional_map.py:49-50)

plus the host-buffer entry (pinned H2D, compute, D2H) and the sharding of pairs across ranks.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

from . import fm as _fm
from . import nn as _nn

__all__ = ["PairBatchHost", "PairBatchDevice", "match_pairs_device", "match_pairs_host", "shard_pairs",
           "gather_results"]


@dataclass
class PairBatchHost:
    """Host-side (numpy, ideally pinned via ``pin()``) ragged batch of mesh pairs."""
    F1: np.ndarray          # [sum n1, d] float32 unit-norm features of the source meshes
    F2: np.ndarray          # [sum n2, d] float32
    off1: np.ndarray        # [P+1] int64
    off2: np.ndarray        # [P+1] int64
    Phi1: Optional[np.ndarray] = None    # [sum n1, K] float64 LBO eigenvectors
    Phi2: Optional[np.ndarray] = None
    evals1: Optional[np.ndarray] = None  # [P, K] float64
    evals2: Optional[np.ndarray] = None
    area1: Optional[np.ndarray] = None   # [sum n1] float64 lumped vertex areas
    area2: Optional[np.ndarray] = None
    _pinned: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False)

    FIELDS = ("F1", "F2", "Phi1", "Phi2", "evals1", "evals2", "area1", "area2")

    @property
    def n_pairs(self):
        return len(self.off1) - 1

    def pin(self):
        """Copies every array into page-locked memory once (so that H2D copies are asynchronous DMA)."""
        for name in self.FIELDS:
            a = getattr(self, name)
            if a is not None and name not in self._pinned:
                t = torch.from_numpy(np.ascontiguousarray(a))
                self._pinned[name] = t.pin_memory() if torch.cuda.is_available() else t
        return self

    def h2d_bytes(self):
        return int(sum(getattr(self, n).nbytes for n in self.FIELDS if getattr(self, n) is not None)
                   + self.off1.nbytes + self.off2.nbytes)

    def to_device(self, device, non_blocking=True):
        get = lambda n: (self._pinned[n] if n in self._pinned else
                         (torch.from_numpy(np.ascontiguousarray(getattr(self, n))) if getattr(self, n) is not None else None))
        t = {n: (get(n).to(device, non_blocking=non_blocking) if get(n) is not None else None) for n in self.FIELDS}
        return PairBatchDevice(off1_h=np.asarray(self.off1, np.int64), off2_h=np.asarray(self.off2, np.int64),
                               device=device, **t)

    def slice_pairs(self, lo, hi):
        """Pairs lo..hi-1 as a new (view-based) batch."""
        r1, r2 = slice(self.off1[lo], self.off1[hi]), slice(self.off2[lo], self.off2[hi])
        g = lambda a, r: None if a is None else a[r]
        return PairBatchHost(F1=self.F1[r1], F2=self.F2[r2], off1=self.off1[lo:hi + 1] - self.off1[lo],
                             off2=self.off2[lo:hi + 1] - self.off2[lo], Phi1=g(self.Phi1, r1), Phi2=g(self.Phi2, r2),
                             evals1=g(self.evals1, slice(lo, hi)), evals2=g(self.evals2, slice(lo, hi)),
                             area1=g(self.area1, r1), area2=g(self.area2, r2))


class PairBatchDevice:
    """The same batch in HBM (layout described in DESIGN.md): row-packed matrices + int64 offsets."""

    def __init__(self, F1, F2, off1_h, off2_h, device, Phi1=None, Phi2=None, evals1=None, evals2=None, area1=None,
                 area2=None):
        self.F1, self.F2, self.Phi1, self.Phi2 = F1, F2, Phi1, Phi2
        self.evals1, self.evals2, self.area1, self.area2 = evals1, evals2, area1, area2
        self.off1_h, self.off2_h = off1_h, off2_h
        self.device = torch.device(device)
        self.off1 = torch.from_numpy(off1_h).to(device, non_blocking=True)
        self.off2 = torch.from_numpy(off2_h).to(device, non_blocking=True)
        self.n_pairs = len(off1_h) - 1
        self.max1 = int(np.diff(off1_h).max()) if self.n_pairs else 0
        self.max2 = int(np.diff(off2_h).max()) if self.n_pairs else 0


def _segment_sum(x, off_dev):
    return torch.segment_reduce(x, "sum", offsets=off_dev)


def fmap_c00(batch: PairBatchDevice):
    """x0[0, 0] = sign(Phi1[0,0] Phi2[0,0]) sqrt(area2 / area1) per pair (pyFM/functional.py:654-658)."""
    a1 = _segment_sum(batch.area1, batch.off1)
    a2 = _segment_sum(batch.area2, batch.off2)
    first1 = batch.Phi1[batch.off1[:-1], 0]
    first2 = batch.Phi2[batch.off2[:-1], 0]
    return torch.sign(first1 * first2) * torch.sqrt(a2 / a1)


def match_pairs_device(batch: PairBatchDevice, k: Optional[int] = None, w_descr: float = 1e4, w_lap: float = 1e3,
                       feature_nn: bool = True, functional_map: bool = True, out_dtype=torch.int32, flags: int = 0):
    """Runs the hot path on a device-resident batch.  Returns a dict of device tensors:
    ``nn_p2p_21`` / ``nn_p2p_12`` (feature NN), ``C`` [P,k,k], ``p2p_21`` / ``p2p_12`` (dense-argmax override,
    what compute_surface_map returns in slots 0/1) and ``p2p_21_adjoint`` / ``p2p_12_adjoint`` (kd-tree-equivalent
    searches, slots 10/11).  Indices are local to each pair."""
    out = {}
    if feature_nn:
        # for each vertex of mesh 2 its nearest feature on mesh 1 (rows), and the reverse (columns)
        (r,), (c,) = _nn.nn_argmax(batch.F2, batch.F1, batch.off2, batch.off1, row_epi=(_nn.COSINE_UNIT,),
                                   col_epi=(_nn.COSINE_UNIT,), max_q=batch.max2, max_db=batch.max1, flags=flags,
                                   out_dtype=out_dtype)
        out["nn_p2p_21"], out["nn_p2p_12"] = r, c
    if functional_map:
        if batch.Phi1 is None:
            raise ValueError("functional_map=True needs eigenbases")
        k = batch.Phi1.shape[1] if k is None else int(k)
        A = _fm.project(batch.Phi1, batch.area1, batch.F1, batch.off1_h, k=k)
        B = _fm.project(batch.Phi2, batch.area2, batch.F2, batch.off2_h, k=k)
        C = _fm.fmap_solve(A, B, batch.evals1[:, :k], batch.evals2[:, :k], fmap_c00(batch), w_descr, w_lap)
        res = _fm.fm_to_p2p(C, batch.Phi1[:, :k], batch.Phi2[:, :k], batch.area1, batch.off1_h, batch.off2_h,
                            flags=flags, out_dtype=out_dtype)
        out.update(C=C, p2p_21=res["dense_21"], p2p_12=res["dense_12"], p2p_21_adjoint=res["p2p_21"],
                   p2p_12_adjoint=res["p2p_12"])
    return out


def match_pairs_host(batch: PairBatchHost, device=None, chunk_pairs: int = 64, **kw):
    """Host buffers in, host (numpy) results out: H2D copies, the device pipeline, D2H copies.  Pairs are
    processed in chunks on two alternating streams so that the copies of chunk i+1 overlap the kernels of
    chunk i (pin the batch first with ``batch.pin()``)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    P = batch.n_pairs
    streams = [torch.cuda.Stream(device), torch.cuda.Stream(device)]
    cur = torch.cuda.current_stream(device)
    parts, events = [], []
    pinned = bool(batch._pinned)
    for ci, lo in enumerate(range(0, P, chunk_pairs)):
        hi = min(P, lo + chunk_pairs)
        sub = batch.slice_pairs(lo, hi)
        if pinned:  # views of the pinned tensors keep the DMA path
            r1, r2 = slice(batch.off1[lo], batch.off1[hi]), slice(batch.off2[lo], batch.off2[hi])
            for n in PairBatchHost.FIELDS:
                if n in batch._pinned:
                    sub._pinned[n] = batch._pinned[n][slice(lo, hi) if n.startswith("evals") else (r1 if n.endswith("1") else r2)]
        s = streams[ci % 2]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            dev = sub.to_device(device)
            res = match_pairs_device(dev, **kw)
            host = {n: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t, non_blocking=True)
                    for n, t in res.items()}
            ev = torch.cuda.Event()
            ev.record(s)
        parts.append(host)
        events.append((ev, dev, res))
    for ev, _, _ in events:
        ev.synchronize()
    out = {n: np.concatenate([p[n].numpy() for p in parts]) for n in parts[0]} if parts else {}
    return out


def shard_pairs(n_pairs: int, rank: int, world: int):
    """Contiguous block of pairs owned by ``rank`` (pairs are independent: no data-path collective)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local: torch.Tensor, counts, group=None):
    """All-gather of ragged per-rank index arrays (the only collective of the path): pads to the largest
    shard, ``all_gather_into_tensor``, then trims.  ``counts`` = number of entries per rank (host ints)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    m = int(max(counts))
    buf = torch.zeros(m, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    allb = torch.empty(world * m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(allb, buf, group=group)
    return torch.cat([allb[r * m: r * m + int(counts[r])] for r in range(world)])
