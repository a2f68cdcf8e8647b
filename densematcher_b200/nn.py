"""Device-level API of the fused similarity + argmax kernels (torch tensors in, torch tensors out).

The reference reaches nearest neighbours through ``knn_query`` (sklearn kd-tree, float64;
densematcher/pyFM/spectral/nn_utils.py:4-38) and a dense ``argmax`` over ``Phi2 C Phi1^T A1``
(densematcher/functional_map.py:49-50).  Both are one primitive here:

    row epilogue:  out[i] = argmax_j  <Y[i], X[j]> * scale[j] + bias[j]
    col epilogue:  out[j] = argmax_i  <Y[i], X[j]> * scale[i] + bias[i]

evaluated by ``dm_nn_argmax_f32`` for a whole ragged batch of mesh pairs in one launch sequence.
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

__all__ = ["Epi", "COSINE_UNIT", "COSINE", "EUCLID", "nn_argmax", "debug_scores", "match_dist", "Workspace",
           "as_offsets", "Offsets"]


@dataclass
class Epi:
    """One argmax epilogue.  ``scale``: None | "invnorm" | float64 tensor; ``bias``: None | "euclid" | tensor."""
    scale: Union[None, str, torch.Tensor] = None
    bias: Union[None, str, torch.Tensor] = None


COSINE_UNIT = Epi()                       # rows already unit norm (model.py:169): plain dot-product argmax
COSINE = Epi(scale="invnorm")             # cosine on arbitrary rows
EUCLID = Epi(bias="euclid")               # Euclidean 1-NN == knn_query (nn_utils.py:4-38)


class Workspace:
    """Grow-only device scratch owned by the caller of the C ABI (one per stream/user)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return self.buf


_default_ws = {}


def default_workspace(device, key="nn"):
    """One scratch buffer per (device, purpose, CUDA stream): calls on different streams may overlap."""
    k = (str(device), key, torch.cuda.current_stream(device).cuda_stream)
    if k not in _default_ws:
        _default_ws[k] = Workspace(device)
    return _default_ws[k]


class Offsets:
    """Row offsets of a ragged batch known on both sides: int64 device tensor + host copy (+ the largest
    segment).  Passing one avoids a host->device copy of the offsets in every call."""

    def __init__(self, dev: torch.Tensor, host: np.ndarray):
        self.dev, self.host = dev, np.ascontiguousarray(np.asarray(host, dtype=np.int64))
        self.max = int(np.diff(self.host).max()) if len(self.host) > 1 else 0

    def __len__(self):
        return len(self.host)


def as_offsets(off, device):
    """-> (int64 device tensor [n+1], host numpy int64 [n+1])."""
    if isinstance(off, Offsets):
        return off.dev, off.host
    off_h = np.ascontiguousarray(np.asarray(off, dtype=np.int64))
    return torch.from_numpy(off_h).to(device, non_blocking=True), off_h


def _fill_epi(e: Epi, out, keep):
    s = _lib.NNEpi()
    if e.scale is None:
        s.scale_mode, s.scale = _lib.SCALE_NONE, None
    elif isinstance(e.scale, str):
        assert e.scale == "invnorm"
        s.scale_mode, s.scale = _lib.SCALE_INVNORM, None
    else:
        t = e.scale.to(dtype=torch.float64).contiguous()
        keep.append(t)
        s.scale_mode, s.scale = _lib.SCALE_ARRAY, t.data_ptr()
    if e.bias is None:
        s.bias_mode, s.bias = _lib.BIAS_NONE, None
    elif isinstance(e.bias, str):
        assert e.bias == "euclid"
        s.bias_mode, s.bias = _lib.BIAS_NEG_HALF_SQNORM, None
    else:
        t = e.bias.to(dtype=torch.float64).contiguous()
        keep.append(t)
        s.bias_mode, s.bias = _lib.BIAS_ARRAY, t.data_ptr()
    s.out = out.data_ptr()
    return s


def nn_argmax(Y: torch.Tensor, X: torch.Tensor, q_off=None, db_off=None, *, row_epi: Sequence[Epi] = (COSINE_UNIT,),
              col_epi: Sequence[Epi] = (), max_q: Optional[int] = None, max_db: Optional[int] = None, flags: int = 0,
              out_dtype=torch.int64, workspace: Optional[Workspace] = None, return_stats: bool = False):
    """Fused score + argmax over a ragged batch of pairs.

    Y [total_q, d], X [total_db, d]: float32 (or both float64) CUDA tensors (row stride arbitrary, unit column stride).
    q_off / db_off: int64 offsets (n_pairs+1), host sequence or CUDA tensor (then pass max_q / max_db);
    None means a single pair.  Returns (row_outputs, col_outputs): lists of index tensors holding LOCAL
    indices (into the pair's database rows for row epilogues, query rows for column epilogues).
    """
    lib = _lib.load()
    if not (Y.is_cuda and X.is_cuda):
        raise ValueError("nn_argmax needs CUDA tensors (there is no CPU path)")
    if Y.dtype != X.dtype or Y.dtype not in (torch.float32, torch.float64):
        raise ValueError("operands must both be float32 or both float64")
    f64 = Y.dtype == torch.float64
    if Y.dim() != 2 or X.dim() != 2 or Y.shape[1] != X.shape[1]:
        raise ValueError(f"shape mismatch: Y {tuple(Y.shape)} vs X {tuple(X.shape)}")
    if Y.stride(1) != 1 and Y.shape[0] > 0:
        Y = Y.contiguous()
    if X.stride(1) != 1 and X.shape[0] > 0:
        X = X.contiguous()
    dev = Y.device
    total_q, d = Y.shape
    total_db = X.shape[0]
    if q_off is None:
        q_off, db_off = [0, total_q], [0, total_db]
    if isinstance(q_off, torch.Tensor) and q_off.is_cuda:
        if max_q is None or max_db is None:
            raise ValueError("device offsets need max_q / max_db")
        qo, do = q_off, db_off
        n_pairs = qo.numel() - 1
    else:
        qo, qh = as_offsets(q_off, dev)
        do, dh = as_offsets(db_off, dev)
        n_pairs = len(qh) - 1
        if len(dh) != len(qh) or qh[0] != 0 or dh[0] != 0 or qh[-1] != total_q or dh[-1] != total_db:
            raise ValueError("offsets do not cover the operand rows")
        if np.any(np.diff(qh) < 0) or np.any(np.diff(dh) < 0):
            raise ValueError("offsets must be non-decreasing")
        max_q = int(np.diff(qh).max()) if n_pairs else 0
        max_db = int(np.diff(dh).max()) if n_pairs else 0
    if out_dtype not in (torch.int64, torch.int32):
        raise ValueError("out_dtype must be int32 or int64")
    if out_dtype == torch.int64:
        flags |= _lib.DM_I64_OUT
    n_row, n_col = len(row_epi), len(col_epi)
    if n_row + n_col == 0 or n_row > 2 or n_col > 2:
        raise ValueError("need between 1 and 2 row and/or column epilogues")
    row_out = [torch.empty(total_q, dtype=out_dtype, device=dev) for _ in range(n_row)]
    col_out = [torch.empty(total_db, dtype=out_dtype, device=dev) for _ in range(n_col)]
    if n_pairs == 0 or total_q == 0 or total_db == 0:
        # nothing to search (or nothing to search in): index outputs are empty / zero, no launch
        if total_db == 0 and total_q > 0 and n_row:
            raise ValueError("empty database: no nearest neighbour exists")
        for t in row_out + col_out:
            t.zero_()
        return (row_out, col_out, (0, 0, 0)) if return_stats else (row_out, col_out)
    keep = []
    RowArr = _lib.NNEpi * max(n_row, 1)
    ColArr = _lib.NNEpi * max(n_col, 1)
    rows = RowArr(*[_fill_epi(e, o, keep) for e, o in zip(row_epi, row_out)])
    cols = ColArr(*[_fill_epi(e, o, keep) for e, o in zip(col_epi, col_out)])
    ws_fn, run_fn = ((lib.dm_nn_f64_workspace_bytes, lib.dm_nn_argmax_f64) if f64 else
                     (lib.dm_nn_workspace_bytes, lib.dm_nn_argmax_f32))
    need = ws_fn(n_pairs, total_q, total_db, max_q, max_db, d, n_row, n_col, flags)
    ws = (workspace or default_workspace(dev)).get(need)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = run_fn(Y.data_ptr(), Y.stride(0), qo.data_ptr(), total_q, max_q, X.data_ptr(), X.stride(0),
                    do.data_ptr(), total_db, max_db, n_pairs, d, rows, n_row, cols, n_col, flags,
                    ws.data_ptr(), ws.numel(), stream)
        _lib.check(rc, "dm_nn_argmax_f64" if f64 else "dm_nn_argmax_f32")
        if return_stats:
            st = (C.c_int64 * 4)()
            _lib.check(lib.dm_nn_read_stats(ws.data_ptr(), st, stream), "dm_nn_read_stats")
            return row_out, col_out, (int(st[0]), int(st[1]), int(st[3]))
    return row_out, col_out


def debug_scores(Y, X, flags=0):
    """The fp32-grade score matrix exactly as the selected engine accumulates it (testing aid)."""
    lib = _lib.load()
    S = torch.empty(Y.shape[0], X.shape[0], dtype=torch.float32, device=Y.device)
    ws = default_workspace(Y.device, "dbg").get(max(256, lib.dm_nn_debug_workspace_bytes(Y.shape[0], X.shape[0], Y.shape[1], flags)))
    with torch.cuda.device(Y.device):
        rc = lib.dm_nn_debug_scores_f32(Y.data_ptr(), Y.stride(0), Y.shape[0], X.data_ptr(), X.stride(0), X.shape[0],
                                        Y.shape[1], S.data_ptr(), S.stride(0), flags, ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream(Y.device).cuda_stream)
    _lib.check(rc, "dm_nn_debug_scores_f32")
    return S


def match_dist(Y, X, idx):
    """float64 Euclidean distance |Y[i] - X[idx[i]]| (global row ids), the `dists` of knn_query."""
    lib = _lib.load()
    out = torch.empty(Y.shape[0], dtype=torch.float64, device=Y.device)
    flags = _lib.DM_I64_OUT if idx.dtype == torch.int64 else 0
    with torch.cuda.device(Y.device):
        rc = lib.dm_match_dist_f32(Y.data_ptr(), Y.stride(0), X.data_ptr(), X.stride(0), idx.data_ptr(), Y.shape[0],
                                   Y.shape[1], out.data_ptr(), flags, torch.cuda.current_stream(Y.device).cuda_stream)
    _lib.check(rc, "dm_match_dist_f32")
    return out


def knn_topk(Y: torch.Tensor, X: torch.Tensor, k: int):
    """(dist [nq, k] float64, idx [nq, k] int64): the k nearest rows of X for every row of Y, ordered by (distance,
    index) -- ``knn_query(..., k > 1)`` of the reference (nn_utils.py:4-38).  float64 CUDA tensors, one pair."""
    lib = _lib.load()
    if not (Y.is_cuda and X.is_cuda):
        raise ValueError("knn_topk needs CUDA tensors (there is no CPU path)")
    Y, X = Y.to(torch.float64).contiguous(), X.to(torch.float64).contiguous()
    nq, d = Y.shape
    ndb = X.shape[0]
    idx = torch.empty(nq, k, dtype=torch.int64, device=Y.device)
    dist = torch.empty(nq, k, dtype=torch.float64, device=Y.device)
    need = lib.dm_knn_workspace_bytes(nq, ndb, d, k)
    ws = default_workspace(Y.device, "knn").get(max(need, 256))
    with torch.cuda.device(Y.device):
        rc = lib.dm_knn_f64(Y.data_ptr(), Y.stride(0), nq, X.data_ptr(), X.stride(0), ndb, d, int(k), idx.data_ptr(),
                            dist.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(Y.device).cuda_stream)
    _lib.check(rc, "dm_knn_f64")
    return dist, idx
