"""Providers and consumers of the eigenbasis on either side of the hot path.

* ``load_operator_cache`` reads the on-disk operator cache the reference's DiffusionNet writes
  (densematcher/diffusion_net/geometry.py:425-570: ``verts, faces, k_eig, frames, mass, evals, evecs`` plus CSR
  triplets of ``L``, ``gradX``, ``gradY``; float32 on disk) and returns a ``TriMesh`` carrying that spectrum, i.e.
  the precomputed input the accelerated path expects.
* ``to_basis`` / ``from_basis`` are DiffusionNet's spectral transforms (diffusion_net/geometry.py:572-598);
  ``to_basis`` is the same Phi^T M F contraction as the descriptor projection and runs on the tcgen05 engine,
  ``from_basis`` on the float64 tensor-core GEMM; ``spectral_diffusion`` chains them around the exp(-lambda t)
  scaling (``LearnedTimeDiffusion.forward``, diffusion_net/layers.py:56-67).
* ``lbo_eigs`` is the device eigenbasis provider (``laplacian_spectrum``, pyFM/mesh/laplacian.py:143-182): Chebyshev-
  filtered subspace iteration, csrc/spectral.cu; ``sym_eig`` its dense symmetric eigensolver.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, fm as _fm
from .nn import default_workspace

__all__ = ["load_operator_cache", "to_basis", "from_basis", "spectral_diffusion", "lbo_eigs", "lbo_eigs_many", "sym_eig", "farthest_point_sampling"]


def load_operator_cache(path, k_eig=None):
    """-> ``TriMesh`` with ``eigenvalues``, ``eigenvectors`` (float64 copies of the cached float32 arrays), lumped
    mass ``A`` and the geometry; ``k_eig`` truncates like the reference (geometry.py:494-495)."""
    from .pyFM.mesh import TriMesh
    with np.load(path, allow_pickle=False) as z:
        if "evecs" not in z or "mass" not in z or "evals" not in z:
            raise ValueError(f"{path}: not a DiffusionNet operator cache (missing evals / evecs / mass)")
        k = int(z["k_eig"]) if k_eig is None else int(k_eig)
        if k > z["evecs"].shape[1]:
            raise ValueError(f"{path}: cache holds {z['evecs'].shape[1]} eigenvectors, {k} requested")
        return TriMesh.from_basis(z["evals"][:k], z["evecs"][:, :k], z["mass"], z["verts"] if "verts" in z else None,
                                  z["faces"] if "faces" in z else None)


def to_basis(values: torch.Tensor, basis: torch.Tensor, massvec: torch.Tensor) -> torch.Tensor:
    """(B,V,D) values, (B,V,K) basis, (B,V) mass -> (B,K,D) spectral coefficients basis^T (mass * values)
    (geometry.py:572-583).  CUDA tensors; one ragged-batched tensor-core contraction for the whole batch."""
    squeeze = values.dim() == 2
    if squeeze:
        values, basis, massvec = values[None], basis[None], massvec[None]
    B, V, D = values.shape
    K = basis.shape[-1]
    off = np.arange(B + 1, dtype=np.int64) * V
    out = _fm.project(basis.reshape(B * V, K), massvec.reshape(B * V), values.reshape(B * V, D).float(), off, k=K)
    out = out.to(values.dtype) if values.dtype in (torch.float32, torch.float64) else out
    return out[0] if squeeze else out


def _batched(values, basis):
    squeeze = basis.dim() == 2
    if squeeze:
        values, basis = values[None], basis[None]
    return values, basis, squeeze


def from_basis(values: torch.Tensor, basis: torch.Tensor) -> torch.Tensor:
    """(B,K,D) coefficients, (B,V,K) basis -> (B,V,D) (geometry.py:586-598), batch dimension optional.  One ragged-
    batched float64 tensor-core GEMM (``dm_from_basis``); the result is cast back to ``values.dtype``."""
    if not basis.is_cuda:
        raise ValueError("from_basis needs CUDA tensors (there is no CPU path)")
    values, basis, squeeze = _batched(values, basis)
    B, V, K = basis.shape
    D = values.shape[-1]
    if values.shape[:2] != (B, K):
        raise ValueError(f"coefficients {tuple(values.shape)} do not match the basis {tuple(basis.shape)}")
    lib = _lib.load()
    dev = basis.device
    Phi = _fm._f64(basis.reshape(B * V, K))
    coef = values.to(torch.float64).contiguous()
    out = torch.empty(B * V, D, dtype=torch.float64, device=dev)
    off = torch.arange(B + 1, dtype=torch.int64, device=dev) * V
    with torch.cuda.device(dev):
        rc = lib.dm_from_basis(coef.data_ptr(), Phi.data_ptr(), Phi.stride(0), off.data_ptr(), V, B, K, D,
                               out.data_ptr(), out.stride(0), _fm._stream(dev))
    _lib.check(rc, "dm_from_basis")
    out = out.reshape(B, V, D)
    out = out.to(values.dtype) if values.dtype in (torch.float32, torch.float16, torch.bfloat16) else out
    return out[0] if squeeze else out


def spectral_diffusion(x: torch.Tensor, mass: torch.Tensor, evals: torch.Tensor, evecs: torch.Tensor,
                       time: torch.Tensor, flags: int = 0) -> torch.Tensor:
    """``LearnedTimeDiffusion.forward(x, L, mass, evals, evecs)`` with method 'spectral' (layers.py:56-67):
    (B,V,C) values, (B,V) mass, (B,K) eigenvalues, (B,V,K) eigenvectors, (C,) diffusion times (clamped at 1e-8 like
    layers.py:46-47) -> (B,V,C) ``from_basis(exp(-evals t) * to_basis(x))``, one C-ABI call for the batch."""
    if not x.is_cuda:
        raise ValueError("spectral_diffusion needs CUDA tensors (there is no CPU path)")
    squeeze = x.dim() == 2
    if squeeze:
        x, mass, evals, evecs = x[None], mass[None], evals[None], evecs[None]
    B, V, C = x.shape
    K = evecs.shape[-1]
    if time.numel() != C:
        raise ValueError("Tensor has wrong shape = {}. Last dim shape should have number of channels = {}".format(
            tuple(x.shape), time.numel()))
    lib = _lib.load()
    dev = x.device
    Phi = _fm._f64(evecs.reshape(B * V, K))
    m = _fm._f64(mass.reshape(B * V))
    ev = evals.to(torch.float64).reshape(B, K).contiguous()
    t = torch.clamp(time.detach().to(torch.float64), min=1e-8).contiguous()
    X = x.reshape(B * V, C).to(torch.float32)
    if X.stride(1) != 1:
        X = X.contiguous()
    out = torch.empty(B * V, C, dtype=torch.float64, device=dev)
    off = torch.arange(B + 1, dtype=torch.int64, device=dev) * V
    need = lib.dm_spectral_diffusion_workspace_bytes(B, B * V, V, K, C)
    ws = default_workspace(dev, "fm").get(max(need, 256))
    with torch.cuda.device(dev):
        rc = lib.dm_spectral_diffusion(Phi.data_ptr(), Phi.stride(0), m.data_ptr(), ev.data_ptr(), X.data_ptr(),
                                       X.stride(0), t.data_ptr(), off.data_ptr(), B * V, V, B, K, C, out.data_ptr(),
                                       out.stride(0), int(flags), ws.data_ptr(), ws.numel(), _fm._stream(dev))
    _lib.check(rc, "dm_spectral_diffusion")
    out = out.reshape(B, V, C).to(x.dtype)
    return out[0] if squeeze else out


def sym_eig(A: torch.Tensor):
    """(B,m,m) or (m,m) symmetric float64 CUDA -> (w ascending, V columns = eigenvectors): ``numpy.linalg.eigh`` on the
    device (Householder + implicit QL, one CTA per matrix; m <= 512)."""
    if not A.is_cuda:
        raise ValueError("sym_eig needs CUDA tensors")
    squeeze = A.dim() == 2
    A3 = (A[None] if squeeze else A).to(torch.float64).contiguous()
    B, m, _ = A3.shape
    lib = _lib.load()
    dev = A.device
    w = torch.empty(B, m, dtype=torch.float64, device=dev)
    V = torch.empty(B, m, m, dtype=torch.float64, device=dev)
    ws = default_workspace(dev, "eig").get(max(lib.dm_sym_eig_workspace_bytes(B, m), 256))
    with torch.cuda.device(dev):
        rc = lib.dm_sym_eig(A3.data_ptr(), m, B, w.data_ptr(), V.data_ptr(), ws.data_ptr(), ws.numel(), _fm._stream(dev))
    _lib.check(rc, "dm_sym_eig")
    return (w[0], V[0]) if squeeze else (w, V)


def lbo_eigs_many(Ws, masses, k, device=None, n_streams=16, **kw):
    """Eigenbases of MANY meshes (a dataset's preprocessing: BASELINE config 5 has 599): ``lbo_eigs`` of every mesh, with up
    to ``n_streams`` of them in flight -- one host thread, CUDA stream and workspace each.  The dense Rayleigh-Ritz solves
    of one mesh occupy a single SM (one CTA), so independent meshes overlap almost perfectly; the C call releases the GIL.
    Returns a list of ``(evals, evecs)`` in input order."""
    import concurrent.futures as cf
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = len(Ws)
    if n == 0:
        return []
    res = [None] * n
    res[0] = lbo_eigs(Ws[0], masses[0], k, device=dev, **kw)   # first call alone: one-time kernel attributes are set here
    if n == 1:
        return res
    workers = max(1, min(int(n_streams), n - 1))
    streams = [torch.cuda.Stream(dev) for _ in range(workers)]
    cur = torch.cuda.current_stream(dev)

    def run(w):
        out = []
        with torch.cuda.device(dev), torch.cuda.stream(streams[w]):
            streams[w].wait_stream(cur)
            for i in range(1 + w, n, workers):
                out.append((i, lbo_eigs(Ws[i], masses[i], k, device=dev, **kw)))
        return out

    with cf.ThreadPoolExecutor(max_workers=workers) as ex:
        for part in ex.map(run, range(workers)):
            for i, r in part:
                res[i] = r
    for s_ in streams:
        cur.wait_stream(s_)
    return res


def lbo_eigs(W, mass, k, device=None, tol=1e-10, max_iter=40, degree=0, return_info=False):
    """The k lowest eigenpairs of ``W phi = lambda diag(mass) phi`` on the GPU -- the device counterpart of
    ``laplacian_spectrum`` (pyFM/mesh/laplacian.py:143-182: ``eigsh(W, k, M=A, sigma=-0.01)``).
    ``W``: scipy sparse (any format) cotangent stiffness, ``mass``: (n,) lumped areas.  Returns float64 CUDA tensors
    ``(evals (k,), evecs (n,k))`` with ``evecs^T diag(mass) evecs = I``; ``return_info`` adds a dict."""
    import scipy.sparse as sp
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    Wc = sp.csr_matrix(W).astype(np.float64)
    Wc.sum_duplicates()
    n = Wc.shape[0]
    if not (0 < k <= n):
        raise ValueError(f"k = {k} eigenpairs requested of a {n}-vertex operator")
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    indptr, indices, vals = up(Wc.indptr, np.int64), up(Wc.indices, np.int32), up(Wc.data, np.float64)
    m = up(np.asarray(mass, dtype=np.float64).ravel(), np.float64)
    evals = torch.empty(k, dtype=torch.float64, device=dev)
    evecs = torch.empty(n, k, dtype=torch.float64, device=dev)
    need = lib.dm_lbo_eigs_workspace_bytes(n, Wc.nnz, k)
    if need == 0:
        raise ValueError("bad eigenproblem size")
    ws = default_workspace(dev, "eig").get(need)
    import ctypes as C
    info = (C.c_int * 4)()
    res = C.c_double(0.0)
    with torch.cuda.device(dev):
        rc = lib.dm_lbo_eigs(indptr.data_ptr(), indices.data_ptr(), vals.data_ptr(), Wc.nnz, m.data_ptr(), n, int(k),
                             float(tol), int(max_iter), int(degree), evals.data_ptr(), evecs.data_ptr(), evecs.stride(0),
                             C.cast(info, C.c_void_p), C.cast(C.pointer(res), C.c_void_p), ws.data_ptr(), ws.numel(),
                             _fm._stream(dev))
    _lib.check(rc, "dm_lbo_eigs")
    d = {"iterations": info[0], "converged": bool(info[1]), "block": info[2], "status": info[3], "residual": res.value}
    if not d["converged"]:
        import warnings
        warnings.warn(f"lbo_eigs: residual {res.value:.2e} after {info[0]} iterations (tol {tol:g})")
    return (evals, evecs, d) if return_info else (evals, evecs)


def farthest_point_sampling(verts, size, first=None, off=None):
    """Euclidean farthest point sampling on the device (``TriMesh.extract_fps(size, geodesic=False)``,
    mesh/trimesh.py:870-876; geometry.py:813-851).  ``verts`` (n,3) or ragged (total,3) with ``off``; ``first``: start
    vertex per mesh (the reference draws it at random: ``None`` does the same).  -> int64 CUDA tensor (size,) or
    (n_meshes, size) of mesh-local indices, identical to the numpy loop for the same start vertex."""
    lib = _lib.load()
    V = verts if torch.is_tensor(verts) else torch.from_numpy(np.ascontiguousarray(verts, dtype=np.float64))
    if not V.is_cuda:
        V = V.cuda()
    V = V.to(torch.float64).contiguous()
    dev = V.device
    single = off is None
    oh = np.array([0, V.shape[0]], dtype=np.int64) if single else np.asarray(off, dtype=np.int64)
    n_m = len(oh) - 1
    sizes = np.diff(oh)
    if size > int(sizes.min()):
        raise ValueError(f"cannot sample {size} points of a mesh with {int(sizes.min())} vertices")
    if first is None:
        first = np.random.default_rng().integers(0, sizes)           # geometry.py:839
    f = torch.from_numpy(np.atleast_1d(np.asarray(first, dtype=np.int64))).to(dev)
    od = torch.from_numpy(oh).to(dev)
    out = torch.empty(n_m, size, dtype=torch.int64, device=dev)
    ws = default_workspace(dev, "eig").get(max(lib.dm_fps_workspace_bytes(V.shape[0]), 256))
    with torch.cuda.device(dev):
        rc = lib.dm_fps(V.data_ptr(), od.data_ptr(), V.shape[0], n_m, f.data_ptr(), int(size), out.data_ptr(),
                        ws.data_ptr(), ws.numel(), _fm._stream(dev))
    _lib.check(rc, "dm_fps")
    return out[0] if single else out
