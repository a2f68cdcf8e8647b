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

from . import _lib
from . import fm as _fm
from . import nn as _nn

__all__ = ["PairBatchHost", "PairBatchDevice", "MeshBankDevice", "MeshBankHost", "match_pairs_device", "match_pairs_host",
           "match_bank_pairs", "match_bank_pairs_host", "intra_category_pairs", "shard_pairs", "gather_results"]


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
    """The same batch in HBM (layout described in DESIGN.md): row-packed matrices + int64 offsets.
    ``off1`` / ``off2`` may be host offset arrays (copied to the device here) or ``nn.Offsets``."""

    def __init__(self, F1, F2, off1_h, off2_h, device, Phi1=None, Phi2=None, evals1=None, evals2=None, area1=None,
                 area2=None):
        self.F1, self.F2, self.Phi1, self.Phi2 = F1, F2, Phi1, Phi2
        self.evals1, self.evals2, self.area1, self.area2 = evals1, evals2, area1, area2
        self.device = torch.device(device)
        mk = lambda o: o if isinstance(o, _nn.Offsets) else _nn.Offsets(
            torch.from_numpy(np.ascontiguousarray(o, dtype=np.int64)).to(device, non_blocking=True), o)
        self.o1, self.o2 = mk(off1_h), mk(off2_h)
        self.off1, self.off2 = self.o1.dev, self.o2.dev
        self.off1_h, self.off2_h = self.o1.host, self.o2.host
        self.n_pairs = len(self.off1_h) - 1
        self.max1, self.max2 = self.o1.max, self.o2.max


def fmap_c00(batch: PairBatchDevice):
    """x0[0, 0] = sign(Phi1[0,0] Phi2[0,0]) sqrt(area2 / area1) per pair (pyFM/functional.py:654-658), ``dm_fmap_c00``."""
    lib = _lib.load()
    dev = batch.device
    P1, P2, a1, a2 = _fm._f64(batch.Phi1), _fm._f64(batch.Phi2), _fm._f64(batch.area1), _fm._f64(batch.area2)
    out = torch.empty(batch.n_pairs, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.dm_fmap_c00(P1.data_ptr(), P1.stride(0), batch.off1.data_ptr(), P2.data_ptr(), P2.stride(0),
                             batch.off2.data_ptr(), a1.data_ptr(), a2.data_ptr(), batch.n_pairs, out.data_ptr(),
                             _fm._stream(dev))
    _lib.check(rc, "dm_fmap_c00")
    return out


def match_pairs_device(batch: PairBatchDevice, k: Optional[int] = None, w_descr: float = 1e4, w_lap: float = 1e3,
                       feature_nn: bool = True, functional_map: bool = True, out_dtype=torch.int32, flags: int = 0,
                       fused: bool = True, check: bool = True):
    """Runs the hot path on a device-resident batch.  Returns a dict of device tensors:
    ``nn_p2p_21`` / ``nn_p2p_12`` (feature NN), ``C`` [P,k,k], ``p2p_21`` / ``p2p_12`` (dense-argmax override,
    what compute_surface_map returns in slots 0/1) and ``p2p_21_adjoint`` / ``p2p_12_adjoint`` (kd-tree-equivalent
    searches, slots 10/11).  Indices are local to each pair."""
    out = {}
    if (feature_nn and functional_map and fused and batch.Phi1 is not None and batch.F1.shape[1] <= 512
            and not (flags & _lib.DM_ENGINE_FFMA) and batch.n_pairs > 0):
        # one library call for the whole path; the projections reuse the feature splits of the NN stage
        k = batch.Phi1.shape[1] if k is None else int(k)
        return _fm.match_pairs(batch.F1, batch.F2, batch.Phi1, batch.Phi2, batch.area1, batch.area2, batch.evals1,
                               batch.evals2, batch.o1, batch.o2, k, w_descr, w_lap, flags=flags, out_dtype=out_dtype,
                               check=check)
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
        A = _fm.project(batch.Phi1, batch.area1, batch.F1, batch.o1, k=k)
        B = _fm.project(batch.Phi2, batch.area2, batch.F2, batch.o2, k=k)
        C = _fm.fmap_solve(A, B, batch.evals1[:, :k], batch.evals2[:, :k], fmap_c00(batch), w_descr, w_lap, check=check)
        res = _fm.fm_to_p2p(C, batch.Phi1[:, :k], batch.Phi2[:, :k], batch.area1, batch.o1, batch.o2,
                            flags=flags, out_dtype=out_dtype)
        out.update(C=C, p2p_21=res["dense_21"], p2p_12=res["dense_12"], p2p_21_adjoint=res["p2p_21"],
                   p2p_12_adjoint=res["p2p_12"])
    return out


def hungarian_pairs(batch: PairBatchDevice, C: torch.Tensor, chunk_pairs: Optional[int] = None, max_bytes: int = 4 << 30):
    """Hungarian assignment of every pair's mapped indicator  Phi2 C Phi1^T A1  (functional_map.py:57,78), maximised,
    for a device-resident batch.  The float64 (n2, n1) matrices (32 MB each at N = 2000) are materialised one chunk of
    pairs at a time -- ``chunk_pairs`` pairs, by default as many as fit in ``max_bytes`` (tall problems need a transposed
    copy as well), at least one -- and each chunk is solved by one ``dm_lap_solve`` launch, one CTA per pair (more
    problems than SMs let two of them share an SM).  Returns a list of ``(row_ind, col_ind)`` numpy pairs identical to
    scipy's.  ``dm_lap_solve`` handles up to 8192 rows / columns per problem."""
    k2, k1 = C.shape[1], C.shape[2]
    n1s, n2s = np.diff(batch.off1_h), np.diff(batch.off2_h)
    if chunk_pairs is None:
        per_pair = float(np.max(n1s * n2s)) * 8.0 * 2.0 if len(n1s) else 1.0
        chunk_pairs = int(max(1, min(256, max_bytes // max(per_pair, 1.0))))
    res = []
    for lo in range(0, batch.n_pairs, chunk_pairs):
        hi = min(lo + chunk_pairs, batch.n_pairs)
        a, b = int(batch.off1_h[lo]), int(batch.off1_h[hi])
        c, d = int(batch.off2_h[lo]), int(batch.off2_h[hi])
        # the chunk's indicators in ONE library call (two ragged GEMMs), no per-pair Python loop
        mats = _fm.mapped_indicators(C[lo:hi], batch.Phi1[a:b, :k1], batch.Phi2[c:d, :k2], batch.area1[a:b],
                                     batch.off1_h[lo:hi + 1] - a, batch.off2_h[lo:hi + 1] - c)
        res.extend(_fm.lap_solve(mats, maximize=True))
    return res


def precise_maps(batch: PairBatchDevice, C: torch.Tensor, faces1, face_off):
    """Barycentric precise map (convert.py:186-231, use_adj=True) of every pair in one ``dm_precise_map`` call:
    ``faces1`` are the packed faces of the meshes 1 (vertex ids local to each mesh), ``face_off`` their offsets.
    Returns (face_match [total_n2], bary [total_n2, 3]) on the device."""
    k2, k1 = C.shape[1], C.shape[2]
    lib = _lib.load()
    Phi2, Cc = _fm._f64(batch.Phi2), C.to(torch.float64).contiguous()
    emb2 = torch.empty((Phi2.shape[0], k1), dtype=torch.float64, device=C.device)
    with torch.cuda.device(C.device):                       # emb2[rows of pair p] = Phi2_p[:, :k2] C[p], one ragged GEMM
        rc = lib.dm_from_basis(Cc.data_ptr(), Phi2.data_ptr(), Phi2.stride(0), batch.off2.data_ptr(), batch.max2,
                               batch.n_pairs, k2, k1, emb2.data_ptr(), emb2.stride(0), _fm._stream(C.device))
    _lib.check(rc, "dm_from_basis")
    emb1 = batch.Phi1[:, :k1].contiguous()
    return _fm.precise_map(emb1, faces1, emb2, batch.off1_h, face_off, batch.off2_h)


class HostStager:
    """Staged host-buffer entry: three streams (H2D, compute, D2H), two input slots in HBM and one pinned
    result buffer per output, all grow-only and reused across calls, so the steady state performs no
    allocation and no host synchronisation until the final wait.  Chunk i+1 is copied in while chunk i
    computes and chunk i-1 is copied out (the two DMA directions are independent engines)."""

    N_SLOTS = 2

    def __init__(self, device):
        self.device = torch.device(device)
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(self.device) for _ in range(3))
        self.slots = [dict(bufs={}, stage=None, in_ev=None, done_ev=None) for _ in range(self.N_SLOTS)]
        self.out_sets = [{}, {}]   # two sets of pinned result buffers, used alternately (see ``run(copy=False)``)
        self.out = self.out_sets[0]
        self.calls = 0

    def _slot_buf(self, slot, name, rows, like):
        """device buffer of the slot for field `name` with at least `rows` rows"""
        b = slot["bufs"].get(name)
        if b is None or b.shape[0] < rows or b.shape[1:] != like.shape[1:] or b.dtype != like.dtype:
            b = torch.empty((int(rows * 1.1) + 1,) + tuple(like.shape[1:]), dtype=like.dtype, device=self.device)
            b.record_stream(self.comp)
            slot["bufs"][name] = b
        return b[:rows]

    def _out_buf(self, name, rows, like):
        b = self.out.get(name)
        if b is None or b.shape[0] < rows or b.shape[1:] != like.shape[1:] or b.dtype != like.dtype:
            b = torch.empty((rows,) + tuple(like.shape[1:]), dtype=like.dtype, pin_memory=True)
            self.out[name] = b
        return b

    def run(self, batch: "PairBatchHost", chunk_pairs: int, copy: bool = True, **kw):
        """``copy=False`` returns numpy VIEWS of the pinned result buffers: no extra host pass over the results, valid
        until the call after the next one on this stager (the two buffer sets alternate)."""
        if not batch._pinned:
            batch.pin()
        self.calls += 1
        self.out = self.out_sets[self.calls & 1]
        P = batch.n_pairs
        cur = torch.cuda.current_stream(self.device)
        for s in (self.h2d, self.comp, self.d2h):
            s.wait_stream(cur)
        off1, off2 = np.asarray(batch.off1, np.int64), np.asarray(batch.off2, np.int64)
        keep, outs, n_status = [], {}, 0
        for ci, lo in enumerate(range(0, P, chunk_pairs)):
            hi = min(P, lo + chunk_pairs)
            slot = self.slots[ci % self.N_SLOTS]
            r1, r2, rp = slice(off1[lo], off1[hi]), slice(off2[lo], off2[hi]), slice(lo, hi)
            n = hi - lo
            if slot["in_ev"] is not None:
                slot["in_ev"].synchronize()       # the pinned offset staging of this slot is free again
            if slot["stage"] is None or slot["stage"].shape[1] < n + 1:
                slot["stage"] = torch.empty(2, chunk_pairs + 1, dtype=torch.int64, pin_memory=True)
            o1h, o2h = off1[lo:hi + 1] - off1[lo], off2[lo:hi + 1] - off2[lo]
            slot["stage"][0, :n + 1] = torch.from_numpy(o1h)
            slot["stage"][1, :n + 1] = torch.from_numpy(o2h)
            with torch.cuda.stream(self.h2d):
                if slot["done_ev"] is not None:
                    self.h2d.wait_event(slot["done_ev"])  # compute of the chunk that used this slot has finished
                dev_t = {}
                for name in PairBatchHost.FIELDS:
                    src = batch._pinned.get(name)
                    if src is None:
                        dev_t[name] = None
                        continue
                    src = src[rp if name.startswith("evals") else (r1 if name.endswith("1") else r2)]
                    dst = self._slot_buf(slot, name, src.shape[0], src)
                    dst.copy_(src, non_blocking=True)
                    dev_t[name] = dst
                offd = self._slot_buf(slot, "_off", 2, slot["stage"])[:, :n + 1]
                offd.copy_(slot["stage"][:, :n + 1], non_blocking=True)
                slot["in_ev"] = torch.cuda.Event()
                slot["in_ev"].record(self.h2d)
            with torch.cuda.stream(self.comp):
                self.comp.wait_event(slot["in_ev"])
                dev = PairBatchDevice(off1_h=_nn.Offsets(offd[0], o1h), off2_h=_nn.Offsets(offd[1], o2h),
                                      device=self.device, **dev_t)
                res = match_pairs_device(dev, check=False, **kw)  # the status words are checked after the final wait
                slot["done_ev"] = torch.cuda.Event()
                slot["done_ev"].record(self.comp)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(slot["done_ev"])
                st = res.pop("status", None)
                if st is not None:
                    sb = self._out_buf("_status", (P + chunk_pairs - 1) // chunk_pairs, st[None])
                    sb[ci].copy_(st, non_blocking=True)
                    n_status = ci + 1
                for name, t in res.items():
                    rows, sl = ((P, rp) if name == "C" else
                                (int(off2[-1]), r2) if "_21" in name else (int(off1[-1]), r1))
                    ob = self._out_buf(name, rows, t)
                    ob[sl].copy_(t, non_blocking=True)
                    outs[name] = rows
            keep.append((res, st))  # results stay alive until the D2H copies have run
        self.d2h.synchronize()
        cur.wait_stream(self.comp)
        if n_status and bool(self.out["_status"][:n_status, 0].any()):
            raise _lib.DMError("match_pairs_host: a functional-map system was not positive definite (rank-deficient "
                               "descriptors or zero weights)")
        if copy:
            return {n: self.out[n][:rows].numpy().copy() for n, rows in outs.items()}
        return {n: self.out[n][:rows].numpy() for n, rows in outs.items()}


_stagers = {}


def match_pairs_host(batch: PairBatchHost, device=None, chunk_pairs: int = 16, copy: bool = True, **kw):
    """Host buffers in, host (numpy) results out -- the call a user of the reference would make per batch:
    pinned H2D copies, the device pipeline, D2H copies, overlapped chunk by chunk (see ``HostStager``).
    ``copy=False``: results are views of pinned buffers that stay valid until the call after the next one."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    st = _stagers.get(str(device))
    if st is None:
        st = _stagers[str(device)] = HostStager(device)
    return st.run(batch, int(chunk_pairs), copy=copy, **kw)


class MeshBankDevice:
    """A dataset of meshes resident in HBM once (features, eigenbasis, eigenvalues, vertex areas, ragged-packed by
    ``off``).  This is the layout for dataset-shaped workloads (BASELINE config 5: 599 meshes, every intra-category
    pair): each mesh is uploaded once (~4.7 MB at N = 2000, d = 384, K = 100) instead of once per pair it takes part in,
    and -- ``match`` -- everything of the per-pair path that depends on one mesh only (operand splits, row norms, the
    projection Phi^T A F) is computed once per mesh (``dm_bank_prepare``); the pairs are two id lists and the kernels read
    each pair's operands through its meshes' rows in the bank (``dm_match_bank_pairs``): nothing is gathered per pair.
    ``assemble`` (a device gather of the pairs' rows into a ``PairBatchDevice``) remains for the stages that want a
    packed batch (ZoomOut / ICP ladders, the dense-map fit) and for configurations the bank kernels do not cover."""

    def __init__(self, F, off, Phi=None, evals=None, area=None, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        to = lambda a, dt: None if a is None else torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a,
                                                                   dtype=dt).to(device)
        self.device = device
        self.F, self.Phi, self.area = to(F, torch.float32), to(Phi, torch.float64), to(area, torch.float64)
        self.evals = to(evals, torch.float64)                     # [n_meshes, K]
        self.off_h = np.ascontiguousarray(np.asarray(off, dtype=np.int64))
        self.off = torch.from_numpy(self.off_h).to(device)
        self.n_meshes = len(self.off_h) - 1
        self.sizes_h = np.diff(self.off_h)
        self._states = {}

    def bank_supported(self, k, flags=0):
        """whether ``dm_match_bank_pairs`` covers this bank (tensor-core engines: d <= 512, k <= 128)"""
        return (self.Phi is not None and self.evals is not None and self.area is not None and self.F.shape[1] <= 512
                and 2 <= int(k) <= 128 and not (flags & _lib.DM_ENGINE_FFMA))

    def prepared(self, k, need=None):
        """``fm.BankState`` for eigenbasis width k: the once-per-mesh preparation, made on first use and kept.
        ``need = (lo, hi)``: only that range of meshes has to be prepared (what a rank's block of pairs touches); a state
        that does not cover it yet is extended by the missing meshes.  Default: the whole bank."""
        k = int(k)
        st = getattr(self, "_states", None)
        if st is None:
            st = self._states = {}
        if k not in st:
            st[k] = _fm.bank_prepare(self.F, self.Phi, self.area, self.evals, _nn.Offsets(self.off, self.off_h), k,
                                     state=getattr(self, "_state_buf", None), mesh_range=need)
        elif not st[k].covers(*(need or (0, self.n_meshes))):
            _fm.bank_prepare(None, None, None, None, None, k, mesh_range=need, bank=st[k])
        return st[k]

    def match(self, src_ids, dst_ids, k=None, w_descr: float = 1e4, w_lap: float = 1e3, out_dtype=torch.int32,
              flags: int = 0, check: bool = True):
        """The hot path for the pairs (src_ids[p] -> mesh 1, dst_ids[p] -> mesh 2) without assembling them: same result
        dict, bit for bit, as ``match_pairs_device(self.assemble(src_ids, dst_ids))``, plus ``off1`` / ``off2`` host
        offsets of the packed outputs."""
        src_ids, dst_ids = np.asarray(src_ids, np.int64), np.asarray(dst_ids, np.int64)
        k = self.Phi.shape[1] if k is None else int(k)
        n = len(src_ids)
        # the meshes these pairs touch must be prepared; a state that already exists is only extended (each mesh is
        # prepared once), a fresh one covers the whole bank unless a driver asked for a range first (match_bank_pairs)
        need = (int(min(src_ids.min(), dst_ids.min())), int(max(src_ids.max(), dst_ids.max())) + 1) if n else None
        bank = self.prepared(k, need if k in getattr(self, "_states", {}) else None)
        o1 = np.concatenate([[0], np.cumsum(self.sizes_h[src_ids])]).astype(np.int64)
        o2 = np.concatenate([[0], np.cumsum(self.sizes_h[dst_ids])]).astype(np.int64)
        # one pinned staging buffer for ids and offsets: no host synchronisation
        stage = torch.empty(4 * n + 2, dtype=torch.int64, pin_memory=True)
        stage[:n] = torch.from_numpy(np.ascontiguousarray(src_ids))
        stage[n:2 * n] = torch.from_numpy(np.ascontiguousarray(dst_ids))
        stage[2 * n:3 * n + 1] = torch.from_numpy(o1)
        stage[3 * n + 1:] = torch.from_numpy(o2)
        dev = stage.to(self.device, non_blocking=True)
        res = _fm.match_bank_pairs(bank, dev[:n], dev[n:2 * n], _nn.Offsets(dev[2 * n:3 * n + 1], o1),
                                   _nn.Offsets(dev[3 * n + 1:], o2), w_descr, w_lap, flags=flags, out_dtype=out_dtype,
                                   check=check)
        res["off1"], res["off2"] = o1, o2
        return res

    def _rows(self, mesh_ids_h):
        """global row indices of the listed meshes, concatenated (device int64) + packed offsets (host, device).
        No host synchronisation: the id list goes up through a pinned staging buffer, every size that the device ops
        need is known on the host."""
        sizes = self.sizes_h[mesh_ids_h]
        o = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        n, total = len(sizes), int(o[-1])
        stage = torch.empty(2 * n + 1, dtype=torch.int64, pin_memory=True)
        stage[:n] = torch.from_numpy(np.ascontiguousarray(mesh_ids_h, dtype=np.int64))
        stage[n:] = torch.from_numpy(o)
        dev = stage.to(self.device, non_blocking=True)
        ids, od = dev[:n], dev[n:]
        start = self.off[ids]                                                    # [P]
        seg = torch.repeat_interleave(torch.arange(n, device=self.device), od[1:] - od[:-1], output_size=total)
        rows = start[seg] + (torch.arange(total, device=self.device) - od[:-1][seg])
        return rows, o, od, ids

    def assemble(self, src_ids, dst_ids):
        """PairBatchDevice of the pairs (src_ids[p] -> mesh 1, dst_ids[p] -> mesh 2), gathered on the device."""
        src_ids, dst_ids = np.asarray(src_ids, np.int64), np.asarray(dst_ids, np.int64)
        r1, o1h, o1d, i1 = self._rows(src_ids)
        r2, o2h, o2d, i2 = self._rows(dst_ids)
        g = lambda t, r: None if t is None else t.index_select(0, r)
        return PairBatchDevice(F1=g(self.F, r1), F2=g(self.F, r2), off1_h=_nn.Offsets(o1d, o1h), off2_h=_nn.Offsets(o2d, o2h),
                               device=self.device, Phi1=g(self.Phi, r1), Phi2=g(self.Phi, r2),
                               evals1=g(self.evals, i1), evals2=g(self.evals, i2), area1=g(self.area, r1),
                               area2=g(self.area, r2))


@dataclass
class MeshBankHost:
    """A dataset of meshes in host memory (numpy, pinned once by ``pin()``): the input of the BANK-shaped host entry
    ``match_bank_pairs_host``.  Same fields as ``MeshBankDevice``.  ``Phi`` may be float32 (the DiffusionNet operator
    cache stores float32 eigenvectors, diffusion_net/geometry.py:539-560): it is then uploaded as float32 -- half the
    bytes -- and widened on the device, which is exact."""
    F: np.ndarray            # [sum n, d] float32
    off: np.ndarray          # [M + 1] int64
    Phi: np.ndarray          # [sum n, K] float64 | float32
    evals: np.ndarray        # [M, K] float64
    area: np.ndarray         # [sum n] float64
    _pinned: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False)

    FIELDS = ("F", "Phi", "evals", "area")

    def pin(self):
        for name in self.FIELDS:
            if name not in self._pinned:
                t = torch.from_numpy(np.ascontiguousarray(getattr(self, name)))
                self._pinned[name] = t.pin_memory() if torch.cuda.is_available() else t
        return self

    def h2d_bytes(self):
        return int(sum(np.asarray(getattr(self, n)).nbytes for n in self.FIELDS) + np.asarray(self.off).nbytes)


class _BankStager:
    """Grow-only device copies of a host bank + pinned result buffers for ``match_bank_pairs_host``."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(self.device) for _ in range(3))
        self.dev, self.out_sets, self.calls = {}, [{}, {}], 0

    def dev_buf(self, name, like):
        b = self.dev.get(name)
        if b is None or b.shape[0] < like.shape[0] or b.shape[1:] != like.shape[1:] or b.dtype != like.dtype:
            b = torch.empty((int(like.shape[0] * 1.05) + 1,) + tuple(like.shape[1:]), dtype=like.dtype, device=self.device)
            b.record_stream(self.comp)
            self.dev[name] = b
        return b[:like.shape[0]]

    def out_buf(self, out, name, rows, like):
        b = out.get(name)
        if b is None or b.shape[0] < rows or b.shape[1:] != like.shape[1:] or b.dtype != like.dtype:
            b = torch.empty((rows,) + tuple(like.shape[1:]), dtype=like.dtype, pin_memory=True)
            out[name] = b
        return b


_bank_stagers = {}


def match_bank_pairs_host(bank: "MeshBankHost", src_ids, dst_ids, device=None, chunk_pairs: int = 128, copy: bool = True,
                          **kw):
    """Host buffers in, host results out, for DATASET-shaped work (BASELINE config 5): the meshes of ``bank`` cross
    PCIe ONCE per call (pinned H2D of features, eigenbasis, eigenvalues, areas), the pairs are two id lists, the
    batches are assembled on the device, and the index maps / C of every chunk are copied back while the next chunk
    computes.  Against the pair-shaped entry (``match_pairs_host``: 9.4 MB per pair at N = 2000, d = 384, K = 100)
    this moves 4.7 MB per MESH.  Returns a dict of numpy arrays packed in pair order (index maps local to each pair)
    plus ``off1`` / ``off2`` (row offsets of the pairs in those arrays)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = _bank_stagers.setdefault(str(device), _BankStager(device))
    if not bank._pinned:
        bank.pin()
    st.calls += 1
    out = st.out_sets[st.calls & 1]
    src_ids, dst_ids = np.asarray(src_ids, np.int64), np.asarray(dst_ids, np.int64)
    sizes = np.diff(np.asarray(bank.off, np.int64))
    off1 = np.concatenate([[0], np.cumsum(sizes[src_ids])]).astype(np.int64)
    off2 = np.concatenate([[0], np.cumsum(sizes[dst_ids])]).astype(np.int64)
    P = len(src_ids)
    cur = torch.cuda.current_stream(device)
    for s_ in (st.h2d, st.comp, st.d2h):
        s_.wait_stream(cur)
    with torch.cuda.stream(st.h2d):
        d = {n: st.dev_buf(n, bank._pinned[n]) for n in MeshBankHost.FIELDS}
        for n in MeshBankHost.FIELDS:
            d[n].copy_(bank._pinned[n], non_blocking=True)
        up = torch.cuda.Event()
        up.record(st.h2d)
    keep, rows_of, statuses = [], {}, []
    with torch.cuda.stream(st.comp):
        st.comp.wait_event(up)
        Phi = d["Phi"] if d["Phi"].dtype == torch.float64 else d["Phi"].to(torch.float64)  # exact widening on the device
        dbank = MeshBankDevice.__new__(MeshBankDevice)
        dbank.device, dbank.F, dbank.Phi, dbank.area, dbank.evals = device, d["F"], Phi, d["area"], d["evals"]
        dbank.off_h = np.ascontiguousarray(np.asarray(bank.off, np.int64))
        dbank.off = torch.from_numpy(dbank.off_h).to(device, non_blocking=True)
        dbank.n_meshes, dbank.sizes_h = len(dbank.off_h) - 1, sizes
        # the once-per-mesh preparation is redone per call (the bank was just uploaded), into a buffer the stager keeps
        dbank._states, dbank._state_buf = {}, st.dev.get("_bank_state")
    for lo in range(0, P, chunk_pairs):
        hi = min(P, lo + chunk_pairs)
        with torch.cuda.stream(st.comp):
            res = _bank_chunk(dbank, src_ids[lo:hi], dst_ids[lo:hi], check=False, **kw)
            done = torch.cuda.Event()
            done.record(st.comp)
        with torch.cuda.stream(st.d2h):
            st.d2h.wait_event(done)
            stt = res.pop("status", None)
            if stt is not None:
                sb = st.out_buf(out, "_status", (P + chunk_pairs - 1) // chunk_pairs, stt[None])
                sb[lo // chunk_pairs].copy_(stt, non_blocking=True)
                statuses.append(lo // chunk_pairs)
            for name, t in res.items():
                rows, sl = ((P, slice(lo, hi)) if name == "C" else
                            (int(off2[-1]), slice(off2[lo], off2[hi])) if "_21" in name else
                            (int(off1[-1]), slice(off1[lo], off1[hi])))
                st.out_buf(out, name, rows, t)[sl].copy_(t, non_blocking=True)
                rows_of[name] = rows
        keep.append((res, stt))
    for bs in dbank._states.values():
        st.dev["_bank_state"] = bs.state
    st.d2h.synchronize()
    cur.wait_stream(st.comp)
    if statuses and bool(out["_status"][: max(statuses) + 1, 0].any()):
        raise _lib.DMError("match_bank_pairs_host: a functional-map system was not positive definite")
    res = {n: (out[n][:r].numpy().copy() if copy else out[n][:r].numpy()) for n, r in rows_of.items()}
    res["off1"], res["off2"] = off1, off2
    return res


def _bank_chunk(bank: MeshBankDevice, src, dst, check: bool = True, use_bank: bool = True, **kw):
    """One chunk of pairs of a device bank: through the bank kernels (``MeshBankDevice.match``: nothing gathered, nothing
    prepared per pair) whenever they cover the request, else assembled into a packed batch for ``match_pairs_device``."""
    k = kw.get("k") or (bank.Phi.shape[1] if bank.Phi is not None else None)
    if (use_bank and len(src) > 0 and kw.get("feature_nn", True) and kw.get("functional_map", True) and kw.get("fused", True)
            and k is not None and bank.bank_supported(k, kw.get("flags", 0))):
        res = bank.match(src, dst, k=k, w_descr=kw.get("w_descr", 1e4), w_lap=kw.get("w_lap", 1e3),
                         out_dtype=kw.get("out_dtype", torch.int32), flags=kw.get("flags", 0), check=check)
        res.pop("off1"), res.pop("off2")
        return res
    return match_pairs_device(bank.assemble(src, dst), check=check, **kw)


def intra_category_pairs(categories):
    """All ordered pairs (i, j), i != j, of meshes that share a category label (BASELINE config 5), grouped by
    category so that consecutive pairs reuse the same meshes (L2 locality).  Returns two int64 arrays."""
    categories = np.asarray(categories)
    src, dst = [], []
    for c in np.unique(categories):
        idx = np.nonzero(categories == c)[0]
        ii, jj = np.meshgrid(idx, idx, indexing="ij")
        keep = ii != jj
        src.append(ii[keep]); dst.append(jj[keep])
    cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.zeros(0, np.int64)
    return cat(src), cat(dst)


def match_bank_pairs(bank: MeshBankDevice, src_ids, dst_ids, chunk_pairs: int = 128, rank: int = 0, world: int = 1,
                     to_host: bool = True, **kw):
    """Runs the hot path over a list of pairs drawn from a device-resident mesh bank, ``chunk_pairs`` at a time.
    With ``world > 1`` only this rank's contiguous block of the pair list (``shard_pairs``) is processed.
    Returns (list of per-chunk result dicts, (lo, hi) of the processed block); index maps are local to each pair
    and packed in pair order."""
    src_ids, dst_ids = np.asarray(src_ids, np.int64), np.asarray(dst_ids, np.int64)
    lo, hi = shard_pairs(len(src_ids), rank, world)
    k = kw.get("k") or (bank.Phi.shape[1] if bank.Phi is not None else None)
    if hi > lo and k is not None and bank.bank_supported(k, kw.get("flags", 0)) and kw.get("functional_map", True):
        # once-per-mesh preparation of the (contiguous) range of meshes this rank's block of pairs touches
        bank.prepared(k, pairs_mesh_range(src_ids, dst_ids, lo, hi))
    out = []
    for a in range(lo, hi, chunk_pairs):
        b = min(hi, a + chunk_pairs)
        res = _bank_chunk(bank, src_ids[a:b], dst_ids[a:b], **kw)
        out.append({n: (t.cpu().numpy() if to_host else t) for n, t in res.items()})
    return out, (lo, hi)


def pairs_mesh_range(src_ids, dst_ids, lo: int, hi: int):
    """(first, last + 1) of the mesh ids that the pairs lo .. hi - 1 touch: the contiguous range of a bank that a rank
    owning that block has to prepare (pairs grouped by category touch ~1 / world of a category-sorted bank)."""
    if hi <= lo:
        return (0, 0)
    a, b = np.asarray(src_ids)[lo:hi], np.asarray(dst_ids)[lo:hi]
    return int(min(a.min(), b.min())), int(max(a.max(), b.max())) + 1


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


class PeerGather:
    """The path's one collective -- the all-gather of every rank's index maps and C -- over NVLink PEER MEMORY with the
    copy engines instead of a NCCL kernel.

    Each rank owns two symmetric receive buffers (``torch.distributed._symmetric_memory``: the same allocation mapped into
    every process of the node).  A step packs its results into a send buffer; on a side stream the rank then writes that
    buffer into ITS slot of every peer's receive buffer -- ``world`` plain device-to-device copies whose destinations live
    in the peers' HBM, executed by the DMA engines over NVLink / NVSwitch, no SM involved -- and joins a device-side
    barrier on the symmetric signal pads, after which every slot of the local receive buffer is complete.  Why not NCCL:
    its all-gather kernel holds 16-32 SMs for the whole transfer and spins on them while ranks are skewed, and the
    kernels of this path run one CTA per SM in waves, so every SM taken away is a longer tail (measured at 8 GPUs with a
    gather of all maps + C every step: 6.7 ms per step with NCCL against 4.9 ms alone, profiles/bench_r2_cfg2a_8gpu_first.json).
    The two receive buffers alternate, so the transfer of step i overlaps the kernels of step i + 1; a buffer is rewritten
    two steps later, and the barrier of the step in between orders those writes after every rank's previous step.
    """

    def __init__(self, n_words: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.world, self.rank, self.m = dist.get_world_size(group), dist.get_rank(group), int(n_words)
        self.device = torch.device(device)
        self.recv = [symm.empty(self.world * self.m, dtype=torch.int32, device=self.device) for _ in range(2)]
        self.hdl = [symm.rendezvous(t, group) for t in self.recv]
        self.peer = [[h.get_buffer(q, (self.world * self.m,), torch.int32, 0) for q in range(self.world)] for h in self.hdl]
        self.stream = torch.cuda.Stream(self.device)
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0

    def gather(self, send: torch.Tensor) -> torch.Tensor:
        """``send``: int32 [n_words] written on the current stream; must stay untouched until ``done[i]`` (returned buffer
        index alternates).  Returns the local receive buffer [world * n_words], valid once ``self.stream`` has passed."""
        i = self.i = self.i ^ 1
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        lo = self.rank * self.m
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for d in range(self.world):                      # start with the own slot, then ring order: spreads the links
                q = (self.rank + d) % self.world
                self.peer[i][q][lo:lo + self.m].copy_(send, non_blocking=True)
            self.hdl[i].barrier(channel=0)
            self.done[i].record(self.stream)
        return self.recv[i]

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
