"""Device-level API of the functional-map stages (torch CUDA tensors in / out, batched over pairs).

Mirrors, stage by stage, what ``FunctionalMapping.fit`` / ``get_p2p`` / ``zoomout_refine`` /
``icp_refine`` do in the reference (densematcher/pyFM/functional.py:201-219, :352-487, :564-617),
but for a ragged batch of mesh pairs at once and without ever leaving the GPU.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .nn import Offsets, Workspace, as_offsets, default_workspace

__all__ = ["project", "fmap_solve", "fm_to_p2p", "mapped_indicator", "mapped_indicators", "p2p_to_fm", "spd_solve", "zoomout", "icp", "polar_factor",
           "match_pairs", "bank_prepare", "match_bank_pairs", "BankState", "dense_energy", "fit_dense", "DENSE_TERMS",
           "PairBatch"]


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _read_status(fn_name, ws, dev):
    """The four status words a solve-type call left at the head of its workspace (synchronises the stream)."""
    import ctypes
    out = (ctypes.c_int * 4)()
    with torch.cuda.device(dev):
        rc = getattr(_lib.load(), fn_name)(ws.data_ptr(), out, _stream(dev))
    _lib.check(rc, fn_name)
    return [int(v) for v in out]


def _raise_if_singular(st, what):
    if st[0]:
        raise _lib.DMError(
            f"{what}: a linear system was not positive definite even in float64 (rank-deficient descriptors / basis, "
            "or zero weights); the reference's optimiser would return a finite map here, this closed-form path cannot")


def _f64(t):
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if t.dim() == 2 and t.stride(1) != 1:
        t = t.contiguous()
    if t.dim() != 2 and not t.is_contiguous():
        t = t.contiguous()
    return t


def _offsets(off, total, dev):
    """-> (device offsets, host offsets, max rows)."""
    if off is None:
        off = [0, total]
    if isinstance(off, Offsets):
        if off.host[0] != 0 or off.host[-1] != total:
            raise ValueError("offsets do not cover the rows")
        return off.dev, off.host, off.max
    od, oh = as_offsets(off, dev)
    if oh[0] != 0 or oh[-1] != total or np.any(np.diff(oh) < 0):
        raise ValueError("offsets do not cover the rows")
    return od, oh, int(np.diff(oh).max()) if len(oh) > 1 else 0


def project(Phi, area, F, off=None, k=None, workspace: Optional[Workspace] = None, flags: int = 0):
    """out[m] = Phi_m[:, :k]^T diag(area_m) F_m -> [n_meshes, k, d] float64.
    (optimize/base_functions.py:526-532; TriMesh.project mesh/trimesh.py:533-556)
    Default: tcgen05 split-bf16 engine (fp32-grade, rel. error ~1e-6); ``flags=_lib.DM_F64_GEMM`` for float64."""
    lib = _lib.load()
    dev = Phi.device
    Phi, area = _f64(Phi), _f64(area)
    if F.dtype != torch.float32:
        F = F.to(torch.float32)
    if F.stride(1) != 1:
        F = F.contiguous()
    total = Phi.shape[0]
    k = Phi.shape[1] if k is None else int(k)
    if k > Phi.shape[1]:
        raise ValueError("not enough eigenvectors")
    if F.shape[0] != total or area.numel() != total:
        raise ValueError(f"Phi has {total} rows but F has {F.shape[0]} and area {area.numel()}")
    d = F.shape[1]
    od, oh, max_n = _offsets(off, total, dev)
    n_m = len(oh) - 1
    out = torch.empty(n_m, k, d, dtype=torch.float64, device=dev)
    need = lib.dm_project_workspace_bytes(n_m, total, max_n, k, d)
    ws = (workspace or default_workspace(dev, "fm")).get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_project_ex(Phi.data_ptr(), Phi.stride(0), area.data_ptr(), F.data_ptr(), F.stride(0),
                               od.data_ptr(), total, max_n, n_m, k, d, out.data_ptr(), int(flags), ws.data_ptr(),
                               ws.numel(), _stream(dev))
    _lib.check(rc, "dm_project_ex")
    return out


def fmap_solve(A, B, evals1, evals2, c00, w_descr, w_lap, workspace: Optional[Workspace] = None, check: bool = True,
               return_status: bool = False):
    """Closed-form minimiser of the descriptor + Laplacian energy, column 0 pinned.
    A [P,k1,d], B [P,k2,d], evals1 [P,k1], evals2 [P,k2], c00 [P] -> C [P,k2,k1] float64.
    (FunctionalMapping.fit pyFM/functional.py:352-487 with only w_descr / w_lap active)
    ``check`` (default) synchronises and raises ``DMError`` when a system was singular; ``return_status`` also
    returns [singular, float64 fallbacks, -, refinement steps]."""
    lib = _lib.load()
    dev = A.device
    A, B = _f64(A).contiguous(), _f64(B).contiguous()
    evals1, evals2, c00 = _f64(evals1).contiguous(), _f64(evals2).contiguous(), _f64(c00).contiguous()
    P, k1, d = A.shape
    k2 = B.shape[1]
    if B.shape[0] != P or B.shape[2] != d or evals1.shape != (P, k1) or evals2.shape != (P, k2) or c00.shape != (P,):
        raise ValueError("shape mismatch in fmap_solve")
    C = torch.empty(P, k2, k1, dtype=torch.float64, device=dev)
    need = lib.dm_fmap_solve_workspace_bytes(P, k1, k2, d)
    ws = (workspace or default_workspace(dev, "fm")).get(need)
    with torch.cuda.device(dev):
        rc = lib.dm_fmap_solve(A.data_ptr(), B.data_ptr(), evals1.data_ptr(), evals2.data_ptr(), c00.data_ptr(),
                               float(w_descr), float(w_lap), P, k1, k2, d, C.data_ptr(), ws.data_ptr(), ws.numel(),
                               _stream(dev))
    _lib.check(rc, "dm_fmap_solve")
    if check or return_status:
        st = _read_status("dm_fmap_solve_read_status", ws, dev)
        if check:
            _raise_if_singular(st, "dm_fmap_solve")
        if return_status:
            return C, st
    return C


def fm_to_p2p(C, Phi1, Phi2, area1=None, off1=None, off2=None, want=("p2p_21", "p2p_12", "dense_21", "dense_12"),
              flags=0, out_dtype=torch.int64, workspace: Optional[Workspace] = None):
    """All index outputs of FM_to_p2p + the dense-argmax override from one score pass.
    C [P,k2,k1]; Phi1 [total_n1, >=k1]; Phi2 [total_n2, >=k2]; area1 [total_n1].
    Returns a dict of LOCAL index tensors.  (pyFM/spectral/convert.py:96-147; functional_map.py:49-50)"""
    lib = _lib.load()
    dev = C.device
    C = _f64(C)
    if C.dim() == 2:
        C = C[None]
    C = C.contiguous()
    Phi1, Phi2 = _f64(Phi1), _f64(Phi2)
    P, k2, k1 = C.shape
    n1, n2 = Phi1.shape[0], Phi2.shape[0]
    if k1 > Phi1.shape[1] or k2 > Phi2.shape[1]:
        raise AssertionError("At least k eigenvectors should be provided")  # convert.py:129-132
    o1, o1h, max1 = _offsets(off1, n1, dev)
    o2, o2h, max2 = _offsets(off2, n2, dev)
    if len(o1h) - 1 != P or len(o2h) - 1 != P:
        raise ValueError("offsets / batch mismatch")
    if out_dtype == torch.int64:
        flags |= _lib.DM_I64_OUT
    outs = {}
    for name in want:
        n = n2 if name.endswith("_21") else n1
        outs[name] = torch.empty(n, dtype=out_dtype, device=dev)
    a1 = _f64(area1).contiguous() if area1 is not None else None
    if "dense_21" in outs and a1 is None:
        raise ValueError("dense_21 needs area1")
    need = lib.dm_fm_to_p2p_workspace_bytes(P, n1, n2, max1, max2, k1, k2, flags)
    ws = (workspace or default_workspace(dev, "fm")).get(need)
    g = lambda nm: outs[nm].data_ptr() if nm in outs else None
    with torch.cuda.device(dev):
        rc = lib.dm_fm_to_p2p(C.data_ptr(), k1, k2, Phi1.data_ptr(), Phi1.stride(0), o1.data_ptr(), n1, max1,
                              Phi2.data_ptr(), Phi2.stride(0), o2.data_ptr(), n2, max2,
                              a1.data_ptr() if a1 is not None else None, P, g("p2p_21"), g("p2p_12"), g("dense_21"),
                              g("dense_12"), flags, ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_fm_to_p2p")
    return outs


def mapped_indicator(C, Phi1, Phi2, area1):
    """Phi2 C Phi1^T A1 for ONE pair, float64 [n2, n1] (convert.py:144) -- API parity only."""
    lib = _lib.load()
    dev = C.device
    C, Phi1, Phi2, area1 = _f64(C).contiguous(), _f64(Phi1), _f64(Phi2), _f64(area1).contiguous()
    k2, k1 = C.shape
    if k1 > Phi1.shape[1] or k2 > Phi2.shape[1]:
        raise AssertionError("At least k eigenvectors should be provided")
    n1, n2 = Phi1.shape[0], Phi2.shape[0]
    MI = torch.empty(n2, n1, dtype=torch.float64, device=dev)
    need = lib.dm_mapped_indicator_workspace_bytes(n1, k2)
    ws = default_workspace(dev, "fm").get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_mapped_indicator(C.data_ptr(), k1, k2, Phi1.data_ptr(), Phi1.stride(0), n1, Phi2.data_ptr(),
                                     Phi2.stride(0), n2, area1.data_ptr(), MI.data_ptr(), MI.stride(0), ws.data_ptr(),
                                     ws.numel(), _stream(dev))
    _lib.check(rc, "dm_mapped_indicator")
    return MI


def mapped_indicators(C, Phi1, Phi2, area1, off1, off2):
    """Phi2 C Phi1^T A1 of every pair of a ragged batch in one call (``dm_mapped_indicators``).  Returns the list of the
    pairs' float64 [n2_p, n1_p] matrices: views into one [total_n2, max_n1] buffer (contiguous when all meshes 1 have one
    size)."""
    lib = _lib.load()
    dev = C.device
    C, Phi1, Phi2, area1 = _f64(C).contiguous(), _f64(Phi1), _f64(Phi2), _f64(area1).contiguous()
    P, k2, k1 = C.shape
    if k1 > Phi1.shape[1] or k2 > Phi2.shape[1]:
        raise AssertionError("At least k eigenvectors should be provided")
    o1, o1h, max1 = _offsets(off1, Phi1.shape[0], dev)
    o2, o2h, max2 = _offsets(off2, Phi2.shape[0], dev)
    if len(o1h) - 1 != P or len(o2h) - 1 != P or area1.numel() != Phi1.shape[0]:
        raise ValueError("mapped_indicators: C, the offsets and the per-vertex arrays describe different batches")
    MI = torch.empty(Phi2.shape[0], max(max1, 1), dtype=torch.float64, device=dev)
    need = lib.dm_mapped_indicators_workspace_bytes(Phi1.shape[0], k2)
    ws = default_workspace(dev, "fm").get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_mapped_indicators(C.data_ptr(), k1, k2, Phi1.data_ptr(), Phi1.stride(0), o1.data_ptr(), Phi1.shape[0], max1,
                                      Phi2.data_ptr(), Phi2.stride(0), o2.data_ptr(), max2, area1.data_ptr(), P, MI.data_ptr(),
                                      MI.stride(0), ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_mapped_indicators")
    return [MI[int(o2h[p]):int(o2h[p + 1]), :int(o1h[p + 1] - o1h[p])] for p in range(P)]


def p2p_to_fm(p2p_21, Phi1, Phi2, area2=None, off1=None, off2=None, k1=None, k2=None,
              workspace: Optional[Workspace] = None, check_indices: bool = True):
    """C[p] = Phi2[:, :k2]^T (area2 * Phi1[p2p_21, :k1]) -> [P,k2,k1] float64 (convert.py:39-48).
    ``area2=None`` gives the un-weighted product Phi2^T Phi1[p] (the normal-equation right-hand side)."""
    lib = _lib.load()
    dev = Phi1.device
    Phi1, Phi2 = _f64(Phi1), _f64(Phi2)
    k1 = Phi1.shape[1] if k1 is None else int(k1)
    k2 = Phi2.shape[1] if k2 is None else int(k2)
    n1, n2 = Phi1.shape[0], Phi2.shape[0]
    o1, o1h, _ = _offsets(off1, n1, dev)
    o2, o2h, max2 = _offsets(off2, n2, dev)
    P = len(o2h) - 1
    if p2p_21.dtype not in (torch.int32, torch.int64):
        p2p_21 = p2p_21.to(torch.int64)
    p2p_21 = p2p_21.contiguous()
    if p2p_21.numel() != n2:
        raise ValueError("p2p_21 must have one entry per target vertex")
    if len(o1h) != len(o2h):
        raise ValueError("off1 and off2 describe different numbers of pairs")
    if check_indices and n2 > 0:
        # indices are local to each pair: 0 <= p < rows of the pair's source mesh (numpy would raise IndexError)
        n1_of = torch.repeat_interleave(torch.as_tensor(np.diff(o1h), device=dev), torch.as_tensor(np.diff(o2h), device=dev))
        if bool(((p2p_21 < 0) | (p2p_21 >= n1_of)).any()):
            raise IndexError("p2p_21 holds an index outside its source mesh")
    flags = _lib.DM_I64_OUT if p2p_21.dtype == torch.int64 else 0
    a2 = _f64(area2).contiguous() if area2 is not None else None
    C = torch.empty(P, k2, k1, dtype=torch.float64, device=dev)
    need = lib.dm_p2p_to_fm_workspace_bytes(P, max2, k1, k2)
    ws = (workspace or default_workspace(dev, "fm")).get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_p2p_to_fm(p2p_21.data_ptr(), Phi1.data_ptr(), Phi1.stride(0), o1.data_ptr(), Phi2.data_ptr(),
                              Phi2.stride(0), o2.data_ptr(), max2, a2.data_ptr() if a2 is not None else None, P, k1, k2,
                              C.data_ptr(), flags, ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_p2p_to_fm")
    return C


def spd_solve(G, B, check: bool = True):
    """X[b] = G[b]^-1 B[b] for symmetric positive definite G [P, n, n], B [P, n, m] (float64): the normal equations of the
    least-squares branch of p2p_to_FM (convert.py:51)."""
    lib = _lib.load()
    G, B = _f64(G), _f64(B)
    if G.dim() == 2:
        G, B = G[None], B[None]
    G, B = G.contiguous(), B.contiguous()
    P, n, m = B.shape
    if G.shape != (P, n, n):
        raise ValueError("shape mismatch in spd_solve")
    dev = G.device
    X = torch.empty_like(B)
    need = lib.dm_spd_solve_workspace_bytes(P, n)
    ws = default_workspace(dev, "fm").get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_spd_solve(G.data_ptr(), B.data_ptr(), n, m, P, X.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_spd_solve")
    if check:
        _raise_if_singular(_read_status("dm_icp_read_status", ws, dev), "dm_spd_solve")
    return X


def zoomout(C0, Phi1, Phi2, area2, nit, step=1, off1=None, off2=None, return_p2p=False, flags=0,
            out_dtype=torch.int64, workspace: Optional[Workspace] = None):
    """ZoomOut ladder with upstream pyFM semantics (pyFM/refine/zoomout.py:7-115).
    C0 [P,k2,k1] -> C [P,k2+nit*s2,k1+nit*s1] (and the final p2p_21, local ids)."""
    lib = _lib.load()
    dev = C0.device
    C0 = _f64(C0)
    if C0.dim() == 2:
        C0 = C0[None]
    C0 = C0.contiguous()
    Phi1, Phi2, area2 = _f64(Phi1), _f64(Phi2), _f64(area2).contiguous()
    try:
        s1, s2 = step
    except TypeError:
        s1 = s2 = step
    s1, s2 = int(s1), int(s2)
    P, k2, k1 = C0.shape
    n1, n2 = Phi1.shape[0], Phi2.shape[0]
    assert k1 + nit * s1 <= Phi1.shape[1], \
        f"Not enough eigenvectors on source : {k1 + nit * s1} are needed when {Phi1.shape[1]} are provided"
    assert k2 + nit * s2 <= Phi2.shape[1], \
        f"Not enough eigenvectors on target : {k2 + nit * s2} are needed when {Phi2.shape[1]} are provided"
    o1, o1h, max1 = _offsets(off1, n1, dev)
    o2, o2h, max2 = _offsets(off2, n2, dev)
    if len(o1h) - 1 != P or len(o2h) - 1 != P:
        raise ValueError("offsets / batch mismatch")
    if out_dtype == torch.int64:
        flags |= _lib.DM_I64_OUT
    C = torch.empty(P, k2 + nit * s2, k1 + nit * s1, dtype=torch.float64, device=dev)
    p2p = torch.empty(n2, dtype=out_dtype, device=dev) if return_p2p else None
    need = lib.dm_zoomout_workspace_bytes(P, n1, n2, max1, max2, k1, k2, nit, s1, s2, flags)
    ws = (workspace or default_workspace(dev, "fm")).get(need)
    with torch.cuda.device(dev):
        rc = lib.dm_zoomout(C0.data_ptr(), k1, k2, nit, s1, s2, Phi1.data_ptr(), Phi1.stride(0), o1.data_ptr(), n1,
                            max1, Phi2.data_ptr(), Phi2.stride(0), o2.data_ptr(), n2, max2, area2.data_ptr(), P,
                            C.data_ptr(), p2p.data_ptr() if p2p is not None else None, flags, ws.data_ptr(), ws.numel(),
                            _stream(dev))
    _lib.check(rc, "dm_zoomout")
    return (C, p2p) if return_p2p else C


def icp(C0, Phi1, Phi2, nit=10, off1=None, off2=None, return_p2p=False, flags=0, out_dtype=torch.int64,
        workspace: Optional[Workspace] = None, check: bool = True):
    """Spectral ICP (pyFM/refine/icp.py:10-107).  C0 [P,k2,k1] -> refined C (and its p2p_21)."""
    lib = _lib.load()
    dev = C0.device
    C0 = _f64(C0)
    if C0.dim() == 2:
        C0 = C0[None]
    C0 = C0.contiguous()
    Phi1, Phi2 = _f64(Phi1), _f64(Phi2)
    P, k2, k1 = C0.shape
    n1, n2 = Phi1.shape[0], Phi2.shape[0]
    if k1 > Phi1.shape[1] or k2 > Phi2.shape[1]:
        raise AssertionError("At least k eigenvectors should be provided")
    o1, o1h, max1 = _offsets(off1, n1, dev)
    o2, o2h, max2 = _offsets(off2, n2, dev)
    if len(o1h) - 1 != P or len(o2h) - 1 != P:
        raise ValueError("offsets / batch mismatch")
    if out_dtype == torch.int64:
        flags |= _lib.DM_I64_OUT
    C = torch.empty(P, k2, k1, dtype=torch.float64, device=dev)
    p2p = torch.empty(n2, dtype=out_dtype, device=dev) if return_p2p else None
    need = lib.dm_icp_workspace_bytes(P, n1, n2, max1, max2, k1, k2, flags)
    ws = (workspace or default_workspace(dev, "fm")).get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_icp(C0.data_ptr(), k1, k2, int(nit), Phi1.data_ptr(), Phi1.stride(0), o1.data_ptr(), n1, max1,
                        Phi2.data_ptr(), Phi2.stride(0), o2.data_ptr(), n2, max2, P, C.data_ptr(),
                        p2p.data_ptr() if p2p is not None else None, flags, ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_icp")
    if check:
        _raise_if_singular(_read_status("dm_icp_read_status", ws, dev), "dm_icp")
    return (C, p2p) if return_p2p else C


def match_pairs(F1, F2, Phi1, Phi2, area1, area2, evals1, evals2, off1, off2, k, w_descr, w_lap, flags=0,
                out_dtype=torch.int32, workspace: Optional[Workspace] = None, check: bool = True):
    """The whole per-pair hot path in ONE library call (``dm_match_pairs``): feature NN both directions, both
    projections (reusing the feature splits of the NN stage), closed-form C, the four FM->p2p index maps.
    Returns a dict like ``pipeline.match_pairs_device``."""
    lib = _lib.load()
    dev = F1.device
    F1 = F1 if (F1.dtype == torch.float32 and F1.stride(1) == 1) else F1.float().contiguous()
    F2 = F2 if (F2.dtype == torch.float32 and F2.stride(1) == 1) else F2.float().contiguous()
    Phi1, Phi2, area1, area2 = _f64(Phi1), _f64(Phi2), _f64(area1).contiguous(), _f64(area2).contiguous()
    k = int(k)
    if k > Phi1.shape[1] or k > Phi2.shape[1]:
        raise AssertionError("At least k eigenvectors should be provided")
    evals1, evals2 = _f64(evals1)[:, :k].contiguous(), _f64(evals2)[:, :k].contiguous()
    n1, n2, d = F1.shape[0], F2.shape[0], F1.shape[1]
    o1, o1h, max1 = _offsets(off1, n1, dev)
    o2, o2h, max2 = _offsets(off2, n2, dev)
    P = len(o1h) - 1
    if len(o2h) - 1 != P:
        raise ValueError("off1 and off2 describe different numbers of pairs")
    if (Phi1.shape[0] != n1 or Phi2.shape[0] != n2 or area1.numel() != n1 or area2.numel() != n2 or F2.shape[1] != d
            or evals1.shape[0] != P or evals2.shape[0] != P):
        raise ValueError("match_pairs: operand shapes do not agree (rows of F / Phi / area per side, one eigenvalue row per pair)")
    if out_dtype == torch.int64:
        flags |= _lib.DM_I64_OUT
    mk = lambda n: torch.empty(n, dtype=out_dtype, device=dev)
    out = dict(nn_p2p_21=mk(n2), nn_p2p_12=mk(n1), C=torch.empty(P, k, k, dtype=torch.float64, device=dev),
               p2p_21=mk(n2), p2p_12=mk(n1), p2p_21_adjoint=mk(n2), p2p_12_adjoint=mk(n1))
    need = lib.dm_match_pairs_workspace_bytes(P, n1, n2, max1, max2, d, k, flags)
    ws = (workspace or default_workspace(dev, "match")).get(need)
    with torch.cuda.device(dev):
        rc = lib.dm_match_pairs(F1.data_ptr(), F1.stride(0), F2.data_ptr(), F2.stride(0), Phi1.data_ptr(), Phi1.stride(0),
                                Phi2.data_ptr(), Phi2.stride(0), area1.data_ptr(), area2.data_ptr(), evals1.data_ptr(),
                                evals2.data_ptr(), o1.data_ptr(), n1, max1, o2.data_ptr(), n2, max2, P, d, k,
                                float(w_descr), float(w_lap), out["nn_p2p_21"].data_ptr(), out["nn_p2p_12"].data_ptr(),
                                out["C"].data_ptr(), out["p2p_21_adjoint"].data_ptr(), out["p2p_12_adjoint"].data_ptr(),
                                out["p2p_21"].data_ptr(), out["p2p_12"].data_ptr(), flags, ws.data_ptr(), ws.numel(),
                                _stream(dev))
    _lib.check(rc, "dm_match_pairs")
    # the solve stage's status words ([singular, float64 fallbacks, -, refinement steps]) lead the workspace: a
    # stream-ordered copy travels with the results, so that a caller that defers the check (check=False; the staged host
    # entry does, to keep its copy/compute overlap) can still make it after its own synchronisation
    if check:
        _raise_if_singular(ws[:16].view(torch.int32).tolist(), "dm_match_pairs")
    else:
        out["status"] = ws[:16].view(torch.int32).clone()
    return out


class BankState:
    """What ``dm_bank_prepare`` leaves in HBM for a bank of meshes: per-mesh operand splits, row norms and projections
    (opaque to Python), together with the matrices they were derived from (the per-pair call re-reads those for its
    float64 re-evaluations).  ``lo`` / ``hi``: the contiguous range of meshes prepared so far."""

    def __init__(self, F, Phi, area, evals, off_dev, off_host, k, state):
        self.F, self.Phi, self.area, self.evals = F, Phi, area, evals
        self.off, self.off_h, self.k, self.state = off_dev, off_host, int(k), state
        self.n_meshes = len(off_host) - 1
        self.sizes_h = np.diff(off_host)
        self.max_n = int(self.sizes_h.max()) if self.n_meshes else 0
        self.lo = self.hi = 0

    def covers(self, lo, hi):
        return lo >= hi or (self.lo <= lo and hi <= self.hi)


def _bank_prepare_range(bank: BankState, lo, hi, workspace=None):
    lib = _lib.load()
    dev = bank.F.device
    N, d = bank.F.shape
    wneed = lib.dm_bank_prepare_workspace_bytes(bank.n_meshes, N, bank.max_n, d, bank.k)
    ws = (workspace or default_workspace(dev, "bank_prep")).get(wneed)
    with torch.cuda.device(dev):
        rc = lib.dm_bank_prepare(bank.F.data_ptr(), bank.F.stride(0), bank.Phi.data_ptr(), bank.Phi.stride(0),
                                 bank.area.data_ptr(), bank.off.data_ptr(), N, bank.max_n, bank.n_meshes, d, bank.k,
                                 int(lo), int(hi), int(bank.off_h[lo]), int(bank.off_h[hi]), bank.state.data_ptr(),
                                 bank.state.numel(), ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_bank_prepare")


def bank_prepare(F, Phi, area, evals, off, k, workspace: Optional[Workspace] = None, state: Optional[torch.Tensor] = None,
                 mesh_range=None, bank: Optional[BankState] = None):
    """Once-per-mesh preparation of a bank of meshes (``dm_bank_prepare``): F [N, d] float32, Phi [N, K >= k] float64,
    area [N], evals [M, K], off [M + 1] row offsets.  Returns a ``BankState`` for ``match_bank_pairs``.
    ``mesh_range = (lo, hi)`` prepares only those meshes (a rank of a sharded job needs the ones its pairs touch); passing a
    previous ``bank`` extends its prepared range to cover ``mesh_range`` -- only the missing meshes are worked on."""
    lib = _lib.load()
    if bank is None:
        dev = F.device
        if dev.type != "cuda":
            raise ValueError("bank_prepare needs CUDA tensors")
        F = F if (F.dtype == torch.float32 and F.stride(1) == 1) else F.float().contiguous()
        Phi, area, evals = _f64(Phi), _f64(area).contiguous(), _f64(evals)
        k = int(k)
        if k > Phi.shape[1] or k > evals.shape[1]:
            raise AssertionError("At least k eigenvectors should be provided")
        N, d = F.shape
        od, oh, _ = _offsets(off, N, dev)
        M = len(oh) - 1
        if Phi.shape[0] != N or area.numel() != N or evals.shape[0] != M:
            raise ValueError("bank_prepare: rows of F / Phi / area, or the number of eigenvalue rows, do not agree")
        need = lib.dm_bank_state_bytes(M, N, d, k)
        if state is None or state.numel() < need:
            state = torch.empty(max(int(need), 256), dtype=torch.uint8, device=dev)
        bank = BankState(F, Phi, area, evals, od, np.asarray(oh, np.int64), k, state)
    lo, hi = (0, bank.n_meshes) if mesh_range is None else (max(0, int(mesh_range[0])), min(bank.n_meshes, int(mesh_range[1])))
    if lo >= hi or bank.covers(lo, hi):
        return bank
    if bank.lo == bank.hi:                      # nothing prepared yet
        _bank_prepare_range(bank, lo, hi, workspace)
        bank.lo, bank.hi = lo, hi
        return bank
    if lo < bank.lo:                            # extend to the left / right: the prepared range stays contiguous
        _bank_prepare_range(bank, lo, bank.lo, workspace)
        bank.lo = lo
    if hi > bank.hi:
        _bank_prepare_range(bank, bank.hi, hi, workspace)
        bank.hi = hi
    return bank


def match_bank_pairs(bank: BankState, ids1, ids2, off1, off2, w_descr, w_lap, flags=0, out_dtype=torch.int32,
                     workspace: Optional[Workspace] = None, check: bool = True):
    """The per-pair hot path for pairs (ids1[p] -> mesh 1, ids2[p] -> mesh 2) drawn from a prepared bank, in one library
    call (``dm_match_bank_pairs``).  ids1 / ids2: device int64 [P]; off1 / off2: the packing of the outputs (``Offsets`` or
    host arrays: cumulative sizes of the pairs' meshes).  Same result dict, bit for bit, as ``match_pairs`` on the
    assembled batch."""
    import ctypes
    lib = _lib.load()
    dev = bank.F.device
    P = int(ids1.numel())
    if int(ids2.numel()) != P or ids1.dtype != torch.int64 or ids2.dtype != torch.int64:
        raise ValueError("match_bank_pairs: ids1 / ids2 must be int64 device tensors of one length")
    if bank.lo >= bank.hi and P > 0:
        raise ValueError("match_bank_pairs: no mesh of the bank has been prepared")
    tot = lambda o: int((o.host if isinstance(o, Offsets) else np.asarray(o))[-1])
    o1, o1h, max1 = _offsets(off1, tot(off1), dev)
    o2, o2h, max2 = _offsets(off2, tot(off2), dev)
    if len(o1h) - 1 != P or len(o2h) - 1 != P:
        raise ValueError("match_bank_pairs: off1 / off2 must describe one entry per pair")
    n1, n2, d, k = int(o1h[-1]), int(o2h[-1]), bank.F.shape[1], bank.k
    if out_dtype == torch.int64:
        flags |= _lib.DM_I64_OUT
    mk = lambda n: torch.empty(n, dtype=out_dtype, device=dev)
    out = dict(nn_p2p_21=mk(n2), nn_p2p_12=mk(n1), C=torch.empty(P, k, k, dtype=torch.float64, device=dev),
               p2p_21=mk(n2), p2p_12=mk(n1), p2p_21_adjoint=mk(n2), p2p_12_adjoint=mk(n1))
    if P == 0:
        return out
    need = lib.dm_match_bank_pairs_workspace_bytes(P, n1, n2, max1, max2, d, k, flags)
    ws = (workspace or default_workspace(dev, "match")).get(need)
    with torch.cuda.device(dev):
        rc = lib.dm_match_bank_pairs(
            bank.state.data_ptr(), bank.state.numel(), bank.F.data_ptr(), bank.F.stride(0), bank.Phi.data_ptr(),
            bank.Phi.stride(0), bank.area.data_ptr(), bank.evals.data_ptr(), bank.evals.stride(0), bank.off.data_ptr(),
            bank.F.shape[0], bank.n_meshes, bank.lo, bank.hi, ids1.data_ptr(), ids2.data_ptr(), o1.data_ptr(), n1, max1, o2.data_ptr(), n2, max2,
            P, d, k, float(w_descr), float(w_lap), out["nn_p2p_21"].data_ptr(), out["nn_p2p_12"].data_ptr(),
            out["C"].data_ptr(), out["p2p_21_adjoint"].data_ptr(), out["p2p_12_adjoint"].data_ptr(),
            out["p2p_21"].data_ptr(), out["p2p_12"].data_ptr(), flags, ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_match_bank_pairs")
    if check:
        st = (ctypes.c_int * 5)()
        with torch.cuda.device(dev):
            rc = lib.dm_match_bank_pairs_read_status(ws.data_ptr(), P, d, k, st, _stream(dev))
        _lib.check(rc, "dm_match_bank_pairs_read_status")
        if st[4]:
            raise ValueError("match_bank_pairs: a mesh id is outside the prepared range of the bank or off1 / off2 do not "
                             "match the meshes' sizes")
        _raise_if_singular([int(v) for v in st[:4]], "dm_match_bank_pairs")
    else:
        out["status"] = ws[:16].view(torch.int32).clone()
    return out


DENSE_TERMS = ("p2p", "stochastic", "ent", "range01", "sumto1")


def bmm_nt(A, B):
    """C[b] = A[b] @ B[b]^T for contiguous float64 batches [P,m,k] x [P,n,k] -> [P,m,n] on the library's own DMMA GEMM
    (``dm_bmm_nt_f64``): the small products of the fit and of the drop-in mirror, without cuBLAS."""
    lib = _lib.load()
    A, B = _f64(A).contiguous(), _f64(B).contiguous()
    P, m, k = A.shape
    n = B.shape[1]
    if B.shape[0] != P or B.shape[2] != k:
        raise ValueError(f"bmm_nt: {tuple(A.shape)} x {tuple(B.shape)}^T")
    out = torch.empty(P, m, n, dtype=torch.float64, device=A.device)
    with torch.cuda.device(A.device):
        rc = lib.dm_bmm_nt_f64(A.data_ptr(), B.data_ptr(), P, m, n, k, out.data_ptr(), _stream(A.device))
    _lib.check(rc, "dm_bmm_nt_f64")
    return out


def dense_energy(C, Phi1, Phi2, area1, weights, off1=None, off2=None, workspace: Optional[Workspace] = None,
                 flags: int = 0):
    """Dense-map energy terms and their gradient (``dm_dense_energy_ex``).  ``weights``: dict over ``DENSE_TERMS``.
    Returns (energies [P, 5] unweighted, in ``DENSE_TERMS`` order; grad [P, k2, k1] of the weighted sum).
    ``flags=_lib.DM_FAST_LOSS``: logarithm / division of the entropy term in float32 (the reference's precision), ~1e-7."""
    lib = _lib.load()
    dev = C.device
    C = _f64(C)
    if C.dim() == 2:
        C = C[None]
    C = C.contiguous()
    Phi1, Phi2, area1 = _f64(Phi1), _f64(Phi2), _f64(area1).contiguous()
    P, k2, k1 = C.shape
    if k1 > Phi1.shape[1] or k2 > Phi2.shape[1]:
        raise AssertionError("At least k eigenvectors should be provided")
    n1, n2 = Phi1.shape[0], Phi2.shape[0]
    o1, o1h, max1 = _offsets(off1, n1, dev)
    o2, o2h, max2 = _offsets(off2, n2, dev)
    if len(o1h) - 1 != P or len(o2h) - 1 != P or area1.numel() != n1:
        raise ValueError("offsets / batch mismatch")
    w = [float(weights.get(t, 0.0)) for t in DENSE_TERMS]
    energy = torch.zeros(P, 5, dtype=torch.float64, device=dev)
    grad = torch.empty(P, k2, k1, dtype=torch.float64, device=dev)
    need = lib.dm_dense_energy_workspace_bytes(P, n1, n2, max1, max2, k1, k2)
    ws = (workspace or default_workspace(dev, "energy")).get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_dense_energy_ex(C.data_ptr(), k1, k2, Phi1.data_ptr(), Phi1.stride(0), o1.data_ptr(), n1, max1,
                                    Phi2.data_ptr(), Phi2.stride(0), o2.data_ptr(), n2, max2, area1.data_ptr(), P, *w,
                                    energy.data_ptr(), grad.data_ptr(), int(flags), ws.data_ptr(), ws.numel(), _stream(dev))
    _lib.check(rc, "dm_dense_energy_ex")
    return energy, grad


def fit_dense(A, B, evals1, evals2, c00, Phi1, Phi2, area1, weights, w_descr, w_lap, off1=None, off2=None,
              maxiter: int = 1000, history: int = 10, gtol: float = 1e-5, ftol: float = 2.220446049250313e-09,
              check_every: int = 2, return_info: bool = False, fast_loss: bool = True):
    """Batched ON-DEVICE fit of the functional map with the dense-map energy terms of the notebook's default
    ``fit_params`` (example.ipynb cell 11: w_ent, w_sumto1, ...; reference: FunctionalMapping.fit functional.py:352-487
    driving scipy L-BFGS-B over energy_func_std / grad_energy_std, optimize/base_functions.py:480-763, one pair at a time
    with a host round trip per callback).

        E(C) = w_descr/2 |C A - B|^2 + w_lap/2 sum C^2 Delta + sum_t weights[t] * term_t(Phi2 C Phi1^T A1),   C[:, 0] pinned

    Here P pairs are minimised together: an L-BFGS iteration (two-loop recursion over `history` pairs, Armijo
    backtracking) is a handful of batched tensor operations on the device plus ONE ``dm_dense_energy`` launch per trial
    point for all pairs; the host only reads a convergence flag every `check_every` iterations.  The stopping rules and
    their defaults are scipy's L-BFGS-B (what the reference runs with): max |g| <= gtol (pgtol = 1e-5) or
    (E_prev - E) / max(|E_prev|, |E|, 1) <= ftol (factr 1e7 * eps); converged pairs are frozen by a mask.
    A [P,k1,d], B [P,k2,d], evals1 [P,k1], evals2 [P,k2], c00 [P]; Phi / area / offsets as in ``dense_energy``.
    ``fast_loss`` (default): the logarithm / division of the entropy term run in float32 (``DM_FAST_LOSS``) -- the
    reference evaluates all of these terms in float32; the energy moves by ~1e-7 relative, far below the stopping rules.
    Returns C [P,k2,k1] float64 (and ``(iterations, energy evaluations)`` with ``return_info``)."""
    dev = A.device
    eflags = _lib.DM_FAST_LOSS if fast_loss else 0
    A, B = _f64(A).contiguous(), _f64(B).contiguous()
    P, k1, _ = A.shape
    k2 = B.shape[1]
    ev1, ev2 = _f64(evals1).reshape(P, k1), _f64(evals2).reshape(P, k2)
    scale = torch.maximum(ev1.max(dim=1).values, ev2.max(dim=1).values)[:, None, None]
    Delta = (ev1[:, None, :] / scale - ev2[:, :, None] / scale) ** 2                      # functional.py:404-405
    AAt = bmm_nt(A, A)                                                                    # [P,k1,k1]
    BAt = bmm_nt(B, A)                                                                    # [P,k2,k1]
    bb = 0.5 * w_descr * (B * B).sum(dim=(1, 2))
    wvec = torch.tensor([float(weights.get(t, 0.0)) for t in DENSE_TERMS], dtype=torch.float64, device=dev)
    free = torch.ones(k1, dtype=torch.float64, device=dev)
    free[0] = 0.0                                                                         # base_functions.py:759
    n_eval = 0

    def energy(C):
        nonlocal n_eval
        n_eval += 1
        CA = bmm_nt(C, AAt)                                                               # AAt is symmetric
        e = 0.5 * w_descr * (C * CA).sum(dim=(1, 2)) - w_descr * (C * BAt).sum(dim=(1, 2)) + bb
        e = e + 0.5 * w_lap * (C * C * Delta).sum(dim=(1, 2))
        g = w_descr * (CA - BAt) + w_lap * (C * Delta)
        ed, gd = dense_energy(C, Phi1, Phi2, area1, weights, off1, off2, flags=eflags)
        return e + (ed * wvec).sum(dim=1), (g + gd) * free

    C = torch.zeros(P, k2, k1, dtype=torch.float64, device=dev)
    C[:, 0, 0] = _f64(c00).reshape(P)                                                     # functional.py:654-658
    f, g = energy(C)
    S, Y, rho = [], [], []
    active = torch.ones(P, dtype=torch.bool, device=dev)
    dot = lambda a, b: (a * b).sum(dim=(1, 2))
    it = 0
    while it < maxiter:
        # two-loop recursion, batched over the pairs
        q = g.clone()
        alphas = []
        for s_, y_, r_ in zip(reversed(S), reversed(Y), reversed(rho)):
            a_ = r_ * dot(s_, q)
            alphas.append(a_)
            q = q - a_[:, None, None] * y_
        if S:
            gamma = dot(S[-1], Y[-1]) / dot(Y[-1], Y[-1]).clamp_min(1e-300)
        else:
            gamma = 1.0 / dot(g, g).sqrt().clamp_min(1e-300)                              # first step: unit length
        q = q * gamma[:, None, None]
        for (s_, y_, r_), a_ in zip(zip(S, Y, rho), reversed(alphas)):
            b_ = r_ * dot(y_, q)
            q = q + (a_ - b_)[:, None, None] * s_
        d = -q
        gd0 = dot(g, d)
        bad_dir = gd0 >= 0                                                                # not a descent direction: steepest descent
        d = torch.where(bad_dir[:, None, None], -g, d)
        gd0 = torch.where(bad_dir, -dot(g, g), gd0)
        # Armijo backtracking (per pair: accepted pairs keep their point while the others shrink the step)
        t = torch.ones(P, dtype=torch.float64, device=dev)
        done = ~active
        Cn, fn, gn = C, f, g
        for _ls in range(30):
            Ct = C + t[:, None, None] * d
            ft, gt = energy(Ct)
            ok = (ft <= f + 1e-4 * t * gd0) & ~done
            Cn = torch.where(ok[:, None, None], Ct, Cn)
            fn = torch.where(ok, ft, fn)
            gn = torch.where(ok[:, None, None], gt, gn)
            done = done | ok
            if bool(done.all()):
                break
            t = torch.where(done, t, 0.5 * t)
        s_ = Cn - C
        y_ = gn - g
        sy = dot(s_, y_)
        good = sy > 1e-30
        if bool(good.any()):
            # pairs whose curvature condition failed contribute a zero update (rho = 0)
            S.append(torch.where(good[:, None, None], s_, torch.zeros_like(s_)))
            Y.append(torch.where(good[:, None, None], y_, torch.zeros_like(y_)))
            rho.append(torch.where(good, 1.0 / sy.clamp_min(1e-300), torch.zeros_like(sy)))
            if len(S) > history:
                S.pop(0), Y.pop(0), rho.pop(0)
        df = (f - fn).abs()
        C, f_prev, f, g = Cn, f, fn, gn
        it += 1
        conv = ((g.abs().amax(dim=(1, 2)) <= gtol) | (df <= ftol * torch.maximum(f_prev.abs(), f.abs()).clamp_min(1.0)) | ~done)
        active = active & ~conv
        if it % check_every == 0 and not bool(active.any()):
            break
    return (C, (it, n_eval)) if return_info else C


def polar_factor(X, flags=0):
    """U I V^T of every matrix of X [P, rows, cols] (float64): the SVD step of ICP (icp.py:39-40)."""
    lib = _lib.load()
    X = _f64(X)
    if X.dim() == 2:
        X = X[None]
    X = X.contiguous()
    P, rows, cols = X.shape
    C = torch.empty_like(X)
    need = lib.dm_polar_factor_workspace_bytes(P, rows, cols)
    ws = default_workspace(X.device, "fm").get(max(need, 256))
    with torch.cuda.device(X.device):
        rc = lib.dm_polar_factor(X.data_ptr(), rows, cols, P, C.data_ptr(), int(flags), ws.data_ptr(), ws.numel(),
                                 _stream(X.device))
    _lib.check(rc, "dm_polar_factor")
    return C


def precise_map(emb1, faces, emb2, off1=None, face_off=None, off2=None, out_dtype=torch.int64):
    """Barycentric precise map (projection_utils.py:16-115): for every row of ``emb2`` the face of the mesh
    (``emb1``, ``faces``) it projects onto and the barycentric coordinates of the projection.  Ragged batches: packed
    rows with ``off1`` / ``face_off`` / ``off2``; face vertex ids are local to their mesh.
    Returns (face_match [n2] (local face index), bary [n2, 3] float64), both on the device."""
    lib = _lib.load()
    emb1, emb2 = _f64(emb1).contiguous(), _f64(emb2).contiguous()
    dev = emb1.device
    faces = torch.as_tensor(faces, device=dev).to(torch.int32).contiguous()
    n1, p = emb1.shape
    n2 = emb2.shape[0]
    if emb2.shape[1] != p:
        raise ValueError("embedding dimensions differ")
    off1_d, off1_h, max_n1 = _offsets(off1, n1, dev)
    off2_d, off2_h, max_n2 = _offsets(off2, n2, dev)
    foff_d, foff_h, _ = _offsets(face_off, faces.shape[0], dev)
    n_pairs = len(off1_h) - 1
    if len(off2_h) - 1 != n_pairs or len(foff_h) - 1 != n_pairs:
        raise ValueError("offset arrays describe different numbers of pairs")
    face_match = torch.empty(n2, dtype=out_dtype, device=dev)
    bary = torch.empty((n2, 3), dtype=torch.float64, device=dev)
    need = lib.dm_precise_map_workspace_bytes(n_pairs, n1, max_n1, n2, faces.shape[0])
    ws = default_workspace(dev, "fm").get(max(need, 256))
    flags = _lib.DM_I64_OUT if out_dtype == torch.int64 else 0
    with torch.cuda.device(dev):
        rc = lib.dm_precise_map(emb1.data_ptr(), emb1.stride(0), off1_d.data_ptr(), n1, max_n1, faces.data_ptr(),
                                foff_d.data_ptr(), faces.shape[0], emb2.data_ptr(), emb2.stride(0), off2_d.data_ptr(), n2,
                                max_n2, n_pairs, p, face_match.data_ptr(), bary.data_ptr(), flags, ws.data_ptr(),
                                ws.numel(), _stream(dev))
    _lib.check(rc, "dm_precise_map")
    return face_match, bary


def lap_solve(costs, maximize=False, return_status=False):
    """``scipy.optimize.linear_sum_assignment`` for a batch of dense float64 matrices resident in HBM (the Hungarian
    slots of ``compute_surface_map``, functional_map.py:57,66,78).  ``costs``: one 2-D CUDA tensor or a list of them
    (shapes may differ).  Returns a list of ``(row_ind, col_ind)`` int64 numpy pairs, identical to scipy's (same
    assignment, same tie-breaking); raises ``ValueError`` with scipy's messages on invalid / infeasible input."""
    lib = _lib.load()
    single = torch.is_tensor(costs)
    mats = [costs] if single else list(costs)
    if not mats:
        return []
    dev = mats[0].device
    mats = [m if (m.dtype == torch.float64 and m.is_contiguous()) else m.to(torch.float64).contiguous() for m in mats]
    for m in mats:
        if m.dim() != 2:
            raise ValueError("expected a matrix (2-D array), got a %d array" % m.dim())
    nr = np.array([m.shape[0] for m in mats], dtype=np.int64)
    nc = np.array([m.shape[1] for m in mats], dtype=np.int64)
    ptr = np.array([m.data_ptr() for m in mats], dtype=np.int64)
    nonempty = nr * nc > 0
    base = int(ptr[nonempty].min()) if nonempty.any() else 0
    n = len(mats)
    meta_h = np.concatenate([np.where(nonempty, (ptr - base) // 8, 0), np.concatenate([[0], np.cumsum(nr)]), np.concatenate([[0], np.cumsum(nc)])])
    meta = torch.from_numpy(meta_h).to(dev)
    cost_off, row_off, col_off = meta[:n], meta[n:2 * n + 1], meta[2 * n + 1:]
    tall = int((nr * nc)[nc < nr].sum())
    out = torch.empty(int(nr.sum()), dtype=torch.int64, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    max_nr, max_nc = int(nr.max()), int(nc.max())
    need = lib.dm_lap_workspace_bytes(n, max_nr, max_nc, tall)
    ws = default_workspace(dev, "lap").get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_lap_solve(base, cost_off.data_ptr(), row_off.data_ptr(), col_off.data_ptr(), n, max_nr, max_nc, tall,
                              int(bool(maximize)), out.data_ptr(), status.data_ptr(), _lib.DM_I64_OUT, ws.data_ptr(),
                              ws.numel(), _stream(dev))
    _lib.check(rc, "dm_lap_solve")
    out_h, st_h = out.cpu().numpy(), status.cpu().numpy()
    if not return_status:
        if (st_h == 1).any():
            raise ValueError("matrix contains invalid numeric entries")
        if (st_h == 2).any():
            raise ValueError("cost matrix is infeasible")
    res = []
    ro = meta_h[n:2 * n + 1]
    for b in range(n):
        col = out_h[ro[b]:ro[b + 1]]
        rows = np.nonzero(col >= 0)[0].astype(np.int64)
        res.append((rows, col[rows].copy()))
    if return_status:
        return res, st_h
    return res[0] if single else res


class PairBatch:
    """A ragged batch of mesh pairs resident in HBM: packed features, eigenbases and offsets.

    This is the data layout of the hot path: every per-vertex array of all pairs is
    stacked row-wise; ``off1`` / ``off2`` delimit the pairs."""

    def __init__(self, F1, F2, off1, off2, Phi1=None, Phi2=None, evals1=None, evals2=None, area1=None, area2=None):
        self.F1, self.F2 = F1, F2
        dev = F1.device
        self.off1, self.off1_h = as_offsets(off1, dev)
        self.off2, self.off2_h = as_offsets(off2, dev)
        self.max1 = int(np.diff(self.off1_h).max())
        self.max2 = int(np.diff(self.off2_h).max())
        self.n_pairs = len(self.off1_h) - 1
        self.Phi1, self.Phi2, self.evals1, self.evals2, self.area1, self.area2 = Phi1, Phi2, evals1, evals2, area1, area2
