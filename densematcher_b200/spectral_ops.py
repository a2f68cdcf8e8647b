"""Providers and consumers of the eigenbasis on either side of the hot path.

* ``load_operator_cache`` reads the on-disk operator cache the reference's DiffusionNet writes
  (densematcher/diffusion_net/geometry.py:425-570: ``verts, faces, k_eig, frames, mass, evals, evecs`` plus CSR
  triplets of ``L``, ``gradX``, ``gradY``; float32 on disk) and returns a ``TriMesh`` carrying that spectrum, i.e.
  the precomputed input the accelerated path expects.
* ``to_basis`` / ``from_basis`` are DiffusionNet's spectral transforms (diffusion_net/geometry.py:572-598);
  ``to_basis`` is the same Phi^T M F contraction as the descriptor projection and runs on the tcgen05 engine.
"""
from __future__ import annotations

import numpy as np
import torch

from . import fm as _fm

__all__ = ["load_operator_cache", "to_basis", "from_basis"]


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


def from_basis(values: torch.Tensor, basis: torch.Tensor) -> torch.Tensor:
    """(K,D) coefficients, (V,K) basis -> (V,D) (geometry.py:586-598): a plain library matmul (cuBLAS), nothing to fuse."""
    return torch.matmul(basis, values.to(basis.dtype))
