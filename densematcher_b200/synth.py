"""Synthetic inputs of the hot path: meshes with a genuine Laplace-Beltrami basis, cheap stand-in bases and the
BASELINE.json feature model, for the bench workloads, the smoke test and the test fixtures.

Input generation only -- nothing here is a checker (the CPU oracle lives in ``oracle/``) and nothing here is on the
compute path.  The hot path takes the eigenbasis as a *precomputed input* (SURVEY.md section 8a rows a7/a11), so these
helpers only have to produce realistic inputs: a subdivided icosahedron, smooth deformations of it, the cotangent
stiffness / lumped mass pair and its low spectrum.  ``oracle/make_goldens.py`` checks the operators against the
reference's ``mesh/laplacian.py`` (:5-42, :88-140, :143-182).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def icosphere(subdivisions):
    """Unit icosphere: 12 / 42 / 162 / 642 / 2562 vertices for 0..4 subdivisions."""
    g = (1.0 + np.sqrt(5.0)) / 2.0
    verts = [(-1, g, 0), (1, g, 0), (-1, -g, 0), (1, -g, 0), (0, -1, g), (0, 1, g),
             (0, -1, -g), (0, 1, -g), (g, 0, -1), (g, 0, 1), (-g, 0, -1), (-g, 0, 1)]
    V = [np.asarray(v, dtype=np.float64) / np.sqrt(1 + g * g) for v in verts]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
         (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        midpoint = {}

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in midpoint:
                m = V[a] + V[b]
                V.append(m / np.linalg.norm(m))
                midpoint[key] = len(V) - 1
            return midpoint[key]

        nxt = []
        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nxt += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        F = nxt
    return np.asarray(V), np.asarray(F, dtype=np.int64)


def deform(V, scale=(1.0, 1.0, 1.0), bump=0.0, phase=(0.0, 0.0)):
    """Anisotropic scaling times an optional smooth radial bump."""
    V = np.asarray(V, dtype=np.float64)
    r = 1.0 + bump * np.sin(3 * V[:, 0] + phase[0]) * np.cos(2 * V[:, 1] + phase[1])
    return V * np.asarray(scale)[None, :] * r[:, None]


def cotan_stiffness(V, F):
    """Sparse cotangent stiffness matrix W (positive semi-definite, rows sum to 0)."""
    n = V.shape[0]
    I, J, S = [], [], []
    for a, b, c in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
        # angle at vertex c is opposite edge (a, b)
        u = V[F[:, a]] - V[F[:, c]]
        w = V[F[:, b]] - V[F[:, c]]
        cot = np.einsum("ij,ij->i", u, w) / np.linalg.norm(np.cross(u, w), axis=1)
        I.append(F[:, a]); J.append(F[:, b]); S.append(0.5 * cot)
    I, J, S = np.concatenate(I), np.concatenate(J), np.concatenate(S)
    W = sp.coo_matrix((np.concatenate([-S, -S, S, S]),
                       (np.concatenate([I, J, I, J]), np.concatenate([J, I, I, J]))), shape=(n, n))
    return W.tocsc()


def lumped_area(V, F):
    """Per-vertex area: one third of the incident triangle areas."""
    fa = 0.5 * np.linalg.norm(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), axis=1)
    a = np.zeros(V.shape[0])
    for c in range(3):
        np.add.at(a, F[:, c], fa / 3.0)
    return a


def lbo_basis(V, F, k):
    """(evals (k,), Phi (n,k), area (n,)): low spectrum of W Phi = lambda A Phi, Phi^T A Phi = I."""
    W = cotan_stiffness(V, F)
    a = lumped_area(V, F)
    # fixed ARPACK start vector + sign convention: the basis (and every fixture minted from it) regenerates bit for bit
    v0 = np.random.default_rng(12345).standard_normal(V.shape[0])
    evals, Phi = spla.eigsh(W, k=max(20, k), M=sp.diags(a).tocsc(), sigma=-0.01, v0=v0)
    order = np.argsort(evals)
    Phi = Phi[:, order][:, :k]
    piv = np.abs(Phi).argmax(axis=0)
    Phi = Phi * np.sign(Phi[piv, np.arange(Phi.shape[1])])[None, :]
    return evals[order][:k], Phi, a


def bandlimited_features(Phi, d, n_band, rng, unit=True, dtype=np.float32):
    """Smooth per-vertex descriptors: noise restricted to the first ``n_band`` eigenvectors."""
    F = Phi[:, :n_band] @ rng.standard_normal((n_band, d))
    if unit:
        F /= np.maximum(np.linalg.norm(F, axis=1, keepdims=True), 1e-5)
    return F.astype(dtype)


def random_unit_features(n, d, rng, dtype=np.float32):
    """Random unit rows: the BASELINE.json synthetic feature model (model.py:169 contract)."""
    F = rng.standard_normal((n, d))
    F /= np.linalg.norm(F, axis=1, keepdims=True)
    return F.astype(dtype)


def synthetic_basis(n, K, rng):
    """Cheap stand-in LBO basis for throughput runs: A-orthonormal
    columns, constant first column, ascending eigenvalues with lambda_0 = 0."""
    a = rng.uniform(0.5, 1.5, size=n) / n
    M = rng.standard_normal((n, K))
    M[:, 0] = 1.0
    Q, R = np.linalg.qr(np.sqrt(a)[:, None] * M)
    Q = Q * np.sign(np.diag(R))[None, :]
    Phi = Q / np.sqrt(a)[:, None]
    evals = np.concatenate([[0.0], np.cumsum(rng.uniform(0.0, 1.0, size=K - 1)) + 0.5])
    return evals, Phi, a
