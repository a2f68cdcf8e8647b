"""Float64 CPU restatement of DenseMatcher's correspondence hot path (numpy/scipy).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  Importers:
``tests/``, ``__graft_entry__.smoke()``, ``bench.py`` (``cpu_baseline`` leg and
``--impl reference``).

Every function cites the reference file:line (relative to the upstream repo
root, JunzheJosephZhu/DenseMatcher @ 96da493) whose arithmetic it restates.

Parity pinning
--------------
The reference ships no golden vectors or runnable tests for this path
(SURVEY.md section 4), so this oracle is pinned against *outputs of the
reference itself*: ``oracle/make_goldens.py`` imports the unmodified reference
from ``/root/reference`` (authoring container only), runs its ``knn_query``,
``FM_to_p2p``, ``p2p_to_FM``, ``icp_refine``, ``FunctionalMapping.fit`` and
``compute_surface_map`` on seeded inputs and stores the results in
``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` replays them against
this file on every CPU test run.

Third-party arithmetic the reference delegates to (not under /root/reference):
scikit-learn ``NearestNeighbors(algorithm='kd_tree')`` (unpinned upstream,
1.9.0 here), ``scipy.linalg.lstsq`` / ``svd`` / ``scipy.optimize.minimize``
(unpinned, 1.18.1 here).  ``knn_query`` below makes the same sklearn call; the
brute-force ``nn_argmax`` formulation is proven equal to it in the golden tests.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

__all__ = [
    "knn_query", "nn_argmax", "knn_bruteforce", "project", "fmap_c00", "ev_sqdiff",
    "fmap_solve_closed_form", "fmap_energy", "fm_to_p2p", "dense_argmax_override",
    "p2p_to_fm", "zoomout_refine", "icp_refine", "surface_map_arrays", "dense_map_energy", "fmap_fit_lbfgs",
    "hungarian", "lap_shortest_augmenting_path", "tri_closest_point", "project_points_to_triangles", "fm_to_precise_map",
]


# --------------------------------------------------------------------------------------
# nearest neighbours
# --------------------------------------------------------------------------------------
def knn_query(X, Y, k=1, return_distance=False, n_jobs=1):
    """For every row of ``Y`` the Euclidean-nearest row(s) of ``X`` via a kd-tree.

    Same third-party call as densematcher/pyFM/spectral/nn_utils.py:28-30
    (leaf_size=40, kd_tree), same squeeze for k == 1 (:32-34) and the same return
    convention (:36-38).
    """
    from sklearn.neighbors import NearestNeighbors

    nbrs = NearestNeighbors(n_neighbors=k, leaf_size=40, algorithm="kd_tree", n_jobs=n_jobs).fit(X)
    d, idx = nbrs.kneighbors(Y)
    if k == 1:
        d, idx = d.squeeze(), idx.squeeze()
    return (d, idx) if return_distance else idx


def nn_argmax(Y, X, col_scale=None, col_bias=None, axis=1, chunk=1024):
    """``argmax`` over ``j`` (axis=1) or ``i`` (axis=0) of ``<Y[i], X[j]> * s[j] + b[j]`` in float64.

    The single primitive behind the three NN flavours of the reference
    (SURVEY.md fact 2): cosine on unit rows (model.py:169 + nn_utils.py:4-38),
    Euclidean 1-NN (s=1, b=-|x|^2/2; nn_utils.py:4-38) and the dense
    ``mapped_indicator`` argmax (functional_map.py:49-50).  Ties go to the lowest
    index (numpy ``argmax``), which is also what exact duplicate rows produce.
    Returns int64.
    """
    Y = np.asarray(Y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    s = None if col_scale is None else np.asarray(col_scale, dtype=np.float64)
    b = None if col_bias is None else np.asarray(col_bias, dtype=np.float64)
    if axis == 1:
        out = np.empty(Y.shape[0], dtype=np.int64)
        for lo in range(0, Y.shape[0], chunk):
            S = Y[lo:lo + chunk] @ X.T
            if s is not None:
                S *= s[None, :]
            if b is not None:
                S += b[None, :]
            out[lo:lo + chunk] = S.argmax(axis=1)
        return out
    S = Y @ X.T
    if s is not None:
        S *= s[None, :]
    if b is not None:
        S += b[None, :]
    return S.argmax(axis=0).astype(np.int64)


def knn_bruteforce(X, Y):
    """1-NN of each row of ``Y`` among rows of ``X`` (Euclidean) without a tree.

    argmin_j |y - x_j|^2 == argmax_j (<y, x_j> - |x_j|^2 / 2); equals ``knn_query``
    (nn_utils.py:4-38) away from exact distance ties.
    """
    X = np.asarray(X, dtype=np.float64)
    return nn_argmax(Y, X, None, -0.5 * np.einsum("ij,ij->i", X, X))


# --------------------------------------------------------------------------------------
# spectral projection and the functional-map solve
# --------------------------------------------------------------------------------------
def project(Phi, area, F, k=None):
    """``Phi[:, :k].T @ (area[:, None] * F)``: coefficients of ``F`` in the LBO basis.

    densematcher/pyFM/mesh/trimesh.py:533-556 (``TriMesh.project``) and the same
    contraction inside the fit, optimize/base_functions.py:526-532, with the lumped
    (diagonal) mass matrix ``A = diag(area)``.
    """
    Phi = np.asarray(Phi, dtype=np.float64)
    if k is not None:
        Phi = Phi[:, :k]
    return Phi.T @ (np.asarray(area, dtype=np.float64)[:, None] * np.asarray(F, dtype=np.float64))


def fmap_c00(Phi1, Phi2, area1, area2):
    """The pinned entry ``C[0, 0]``: sign(Phi1[0,0] * Phi2[0,0]) * sqrt(area2 / area1).

    densematcher/pyFM/functional.py:654-658 (``get_x0``); ``area`` is the sum of
    the lumped vertex areas (mesh/trimesh.py:206-221).
    """
    sgn = np.sign(Phi1[0, 0] * Phi2[0, 0])
    return float(sgn * np.sqrt(np.sum(area2) / np.sum(area1)))


def ev_sqdiff(evals1, evals2):
    """(k2, k1) squared differences of eigenvalues scaled by the largest one.

    densematcher/pyFM/functional.py:403-405.
    """
    evals1 = np.asarray(evals1, dtype=np.float64)
    evals2 = np.asarray(evals2, dtype=np.float64)
    scale = max(evals1.max(), evals2.max())
    return np.square(evals1[None, :] / scale - evals2[:, None] / scale)


def fmap_energy(C, A, B, Delta, w_descr, w_lap):
    """w_descr * 1/2 |C A - B|^2 + w_lap * 1/2 sum C^2 * Delta.

    optimize/base_functions.py:31-56 (descriptor preservation) and :79-102
    (Laplacian commutativity), weighted as in ``energy_func_std`` :536-544.
    """
    r = C @ A - B
    return 0.5 * w_descr * float(np.sum(r * r)) + 0.5 * w_lap * float(np.sum(C * C * Delta))


def fmap_solve_closed_form(A, B, evals1, evals2, c00, w_descr, w_lap):
    """Exact minimiser of the reference's descriptor + Laplacian energy.

    The reference minimises ``fmap_energy`` with L-BFGS-B from ``x0`` (zeros except
    ``x0[0, 0] = c00``) while zeroing the gradient of column 0
    (functional.py:441,477; base_functions.py:759), i.e. column 0 stays
    ``c00 * e_0``.  With only those two terms the rows of ``C`` decouple
    (SURVEY.md App. A.3):

        c_i[1:] (w_d Abar Abar^T + w_l diag(Delta[i, 1:])) = w_d (B_i - C[i,0] A_0) Abar^T

    ``A``: (k1, d), ``B``: (k2, d).  Returns (k2, k1) float64.
    """
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    k1, k2 = A.shape[0], B.shape[0]
    Delta = ev_sqdiff(np.asarray(evals1)[:k1], np.asarray(evals2)[:k2])
    C = np.zeros((k2, k1))
    C[0, 0] = c00
    Abar = A[1:]
    G = w_descr * (Abar @ Abar.T)
    R = w_descr * ((B - C[:, :1] * A[:1]) @ Abar.T)  # (k2, k1-1)
    for i in range(k2):
        M = G + w_lap * np.diag(Delta[i, 1:])
        C[i, 1:] = scipy.linalg.solve(M, R[i], assume_a="pos")
    return C


# --------------------------------------------------------------------------------------
# functional map <-> point-to-point map
# --------------------------------------------------------------------------------------
def fm_to_p2p(C, Phi1, Phi2, a1=None, nn="brute", want_indicator=True):
    """(p2p_21, p2p_12, mapped_indicator) as the *modified* reference FM_to_p2p.

    densematcher/pyFM/spectral/convert.py:126-147:
      p2p_12 = knn(tree = Phi2[:, :k2] @ C, query = Phi1[:, :k1])          (:134-136)
      p2p_21 = knn(tree = Phi1[:, :k1] @ C.T, query = Phi2[:, :k2])        (:138-140)
      mapped_indicator = Phi2 @ C @ Phi1.T @ A1                            (:144)
    ``use_adj`` is ignored upstream, so it is not a parameter here.  ``a1`` is the
    diagonal of A1.  ``nn='tree'`` uses the sklearn kd-tree exactly as the
    reference; ``'brute'`` the float64 argmax form (identical results away from
    exact ties, 25x faster).
    """
    C = np.asarray(C, dtype=np.float64)
    k2, k1 = C.shape
    assert k1 <= Phi1.shape[1] and k2 <= Phi2.shape[1]
    P1 = np.asarray(Phi1, dtype=np.float64)[:, :k1]
    P2 = np.asarray(Phi2, dtype=np.float64)[:, :k2]
    find = knn_query if nn == "tree" else knn_bruteforce
    emb2 = P2 @ C
    p2p_12 = find(emb2, P1)
    emb1 = P1 @ C.T
    p2p_21 = find(emb1, P2)
    MI = None
    if want_indicator:
        MI = (emb2 @ P1.T) * np.asarray(a1, dtype=np.float64)[None, :]
    return np.asarray(p2p_21, dtype=np.int64), np.asarray(p2p_12, dtype=np.int64), MI


def dense_argmax_override(MI, eta=None):
    """(argmax over axis 1, argmax over axis 0) of ``mapped_indicator * eta[:, None]``.

    densematcher/functional_map.py:49-50 and :76-77; ``eta`` is all ones
    (functional.py:483).
    """
    if eta is not None:
        MI = MI * np.asarray(eta)[:, None]
    return MI.argmax(axis=1).astype(np.int64), MI.argmax(axis=0).astype(np.int64)


def p2p_to_fm(p2p_21, Phi1, Phi2, A2=None):
    """Functional map induced by a vertex map.

    densematcher/pyFM/spectral/convert.py:39-51: pull back ``Phi1[p2p_21]`` (or
    ``P @ Phi1`` for a matrix map), then ``Phi2.T @ (A2 .)`` when the target mass
    is given (1-D, sparse or dense), else least squares ``lstsq(Phi2, pullback)``.
    """
    Phi1 = np.asarray(Phi1, dtype=np.float64)
    Phi2 = np.asarray(Phi2, dtype=np.float64)
    pb = Phi1[np.asarray(p2p_21), :] if np.asarray(p2p_21).ndim == 1 else p2p_21 @ Phi1
    if A2 is None:
        return scipy.linalg.lstsq(Phi2, pb)[0]
    if A2.shape[0] != Phi2.shape[0]:
        raise ValueError("Can't compute exact pseudo inverse with subsampled eigenvectors")
    if getattr(A2, "ndim", 2) == 1:
        return Phi2.T @ (np.asarray(A2)[:, None] * pb)
    return Phi2.T @ (A2 @ pb)


def _steps(step):
    try:
        s1, s2 = step
    except TypeError:
        s1 = s2 = step
    return int(s1), int(s2)


def zoomout_refine(FM_12, Phi1, Phi2, nit=10, step=1, A2=None, subsample=None,
                   return_p2p=False, nn="brute"):
    """ZoomOut with *upstream pyFM* semantics.

    The shipped call ``spectral.FM_to_p2p(FM_12, evects1, evects2, n_jobs=...)``
    (densematcher/pyFM/refine/zoomout.py:40,112) no longer matches the modified
    ``FM_to_p2p`` signature (convert.py:96) and raises TypeError (SURVEY.md fact
    3), so the loop structure follows zoomout.py:7-115 while the conversion is the
    un-modified upstream one it was written against:
        p2p_21 = knn(tree = Phi1[:, :k1] @ C.T, query = Phi2[:, :k2])
    then ``C <- p2p_to_FM(p2p_21, Phi1[:, :k1+s1], Phi2[:, :k2+s2], A2)`` (:42).
    With ``subsample=(sub1, sub2)`` rows are restricted and the map is solved by
    least squares (:104-105).
    """
    C = np.array(FM_12, dtype=np.float64, copy=True)
    s1, s2 = _steps(step)
    k2_0, k1_0 = C.shape
    assert k1_0 + nit * s1 <= Phi1.shape[1], "Not enough eigenvectors on source"
    assert k2_0 + nit * s2 <= Phi2.shape[1], "Not enough eigenvectors on target"
    find = knn_query if nn == "tree" else knn_bruteforce
    E1, E2, area = Phi1, Phi2, A2
    if subsample is not None:
        E1, E2, area = Phi1[subsample[0]], Phi2[subsample[1]], None
    for _ in range(nit):
        k2, k1 = C.shape
        p = find(E1[:, :k1] @ C.T, E2[:, :k2])
        C = p2p_to_fm(p, E1[:, :k1 + s1], E2[:, :k2 + s2], A2=area)
    if return_p2p:
        k2, k1 = C.shape
        return C, np.asarray(find(Phi1[:, :k1] @ C.T, Phi2[:, :k2]), dtype=np.int64)
    return C


def fps_euclidean(V, size, first):
    """Euclidean farthest point sampling from a given start vertex: ``TriMesh.extract_fps(size, geodesic=False)``
    (densematcher/pyFM/mesh/trimesh.py:870-876) -> ``farthest_point_sampling_call`` (mesh/geometry.py:813-851), whose
    random start vertex (:839) is the ``first`` argument here."""
    V = np.asarray(V, dtype=np.float64)
    dist_func = lambda i: np.linalg.norm(V - V[i, None, :], axis=1)     # trimesh.py:871-872
    inds = [int(first)]
    dists = dist_func(inds[0])
    for _ in range(size - 1):                                           # geometry.py:842-848
        newid = int(np.argmax(dists))
        inds.append(newid)
        dists = np.minimum(dists, dist_func(newid))
    return np.asarray(inds, dtype=np.int64)


def mesh_zoomout_refine_p2p(p2p_21, Phi1, Phi2, A2, k_init, nit=10, step=1, subsample=None, p2p_on_sub=False,
                            return_p2p=False):
    """densematcher/pyFM/refine/zoomout.py:164-217 on bare arrays: initial map from the vertex map (on the samples when
    ``p2p_on_sub``, else on all the vertices: :208-211 -> convert.py:89-93), then ``zoomout_refine``."""
    k1, k2 = (k_init, k_init) if np.issubdtype(type(k_init), np.integer) else k_init
    if subsample is not None and p2p_on_sub:
        C0 = p2p_to_fm(p2p_21, Phi1[subsample[0], :k1], Phi2[subsample[1], :k2], A2=None)
    else:
        C0 = p2p_to_fm(p2p_21, Phi1[:, :k1], Phi2[:, :k2], A2=A2)
    return zoomout_refine(C0, Phi1, Phi2, nit=nit, step=step, A2=A2, subsample=subsample, return_p2p=return_p2p)


def icp_refine(FM_12, Phi1, Phi2, nit=10, tol=1e-10, return_p2p=False, nn="brute"):
    """Spectral ICP.

    densematcher/pyFM/refine/icp.py:36-40 per iteration: p2p_21 from
    ``FM_to_p2p`` (the knn of :138-140 in convert.py), ``C = lstsq(Phi2[:, :k2],
    Phi1[p2p_21, :k1])``, then the nearest (partial) isometry ``U @ eye(k2, k1) @
    Vt`` of its SVD.  ``nit`` in (None, 0) iterates until the max-abs change is
    <= ``tol`` (icp.py:84-94), at most 10000 times (:82).
    """
    C = np.array(FM_12, dtype=np.float64, copy=True)
    k2, k1 = C.shape
    P1 = np.asarray(Phi1, dtype=np.float64)[:, :k1]
    P2 = np.asarray(Phi2, dtype=np.float64)[:, :k2]
    find = knn_query if nn == "tree" else knn_bruteforce
    fixed = nit is not None and nit > 0
    for _ in range(nit if fixed else 10000):
        p = find(P1 @ C.T, P2)
        U, _, Vt = scipy.linalg.svd(scipy.linalg.lstsq(P2, P1[p])[0])
        C_new = U @ np.eye(k2, k1) @ Vt
        done = (not fixed) and np.max(np.abs(C - C_new)) <= tol
        C = C_new
        if done:
            break
    if return_p2p:
        return C, np.asarray(find(P1 @ C.T, P2), dtype=np.int64)
    return C


# --------------------------------------------------------------------------------------
# dense-map energy terms and the L-BFGS-B fit
# --------------------------------------------------------------------------------------
DENSE_TERMS = ("p2p", "stochastic", "ent", "range01", "sumto1")


def dense_map_energy(C, Phi1, Phi2, a1, weights):
    """Energy and gradient (w.r.t. ``C``) of the terms defined on the dense map
    ``M = Phi2 @ C @ Phi1.T @ diag(a1)`` (n2, n1), float64, analytic gradients.

    optimize/base_functions.py: ``p2p`` :296-325 sum (M^2 - M)^2; ``doubly_stochastic`` :327-361
    sum_j (sum_i M^2 - n2/n1)^2 + sum_i (sum_j M^2 - 1)^2; ``entropy`` :363-372
    sum -clamp(M,0,1) log(clamp(M,0,1) + 1e-10); ``range01`` :374-385 sum relu(-M)^2 + relu(M-1)^2;
    ``sumto1`` :387-428 (v is None branch) sum_j (colsum - mean)^2 + sum_i (rowsum - mean)^2.
    ``weights``: dict with keys from ``DENSE_TERMS`` (missing = 0).  Returns (energy, grad (k2, k1), per-term dict).
    """
    C = np.asarray(C, dtype=np.float64)
    k2, k1 = C.shape
    P1, P2 = np.asarray(Phi1, np.float64)[:, :k1], np.asarray(Phi2, np.float64)[:, :k2]
    a1 = np.asarray(a1, np.float64)
    n1, n2 = P1.shape[0], P2.shape[0]
    M = (P2 @ C @ P1.T) * a1[None, :]
    G = np.zeros_like(M)                      # dE/dM
    E, parts = 0.0, {}

    def add(name, e, g):
        nonlocal E, G
        w = float(weights.get(name, 0.0))
        parts[name] = e
        if w != 0.0:
            E += w * e
            G += w * g

    if weights.get("p2p", 0):
        q = M * M - M
        add("p2p", float(np.sum(q * q)), 2.0 * q * (2.0 * M - 1.0))
    if weights.get("stochastic", 0):
        M2 = M * M
        cs, rs = M2.sum(0) - n2 / n1, M2.sum(1) - 1.0
        add("stochastic", float(np.sum(cs * cs) + np.sum(rs * rs)), 2.0 * M * (2.0 * cs[None, :] + 2.0 * rs[:, None]))
    if weights.get("ent", 0):
        Mc = np.clip(M, 0.0, 1.0)
        inside = (M >= 0.0) & (M <= 1.0)
        e = float(np.sum(-Mc * np.log(Mc + 1e-10)))
        add("ent", e, np.where(inside, -np.log(Mc + 1e-10) - Mc / (Mc + 1e-10), 0.0))
    if weights.get("range01", 0):
        lo, hi = np.maximum(-M, 0.0), np.maximum(M - 1.0, 0.0)
        add("range01", float(np.sum(lo * lo) + np.sum(hi * hi)), -2.0 * lo + 2.0 * hi)
    if weights.get("sumto1", 0):
        c, r = M.sum(0), M.sum(1)
        dc, dr = c - c.mean(), r - r.mean()
        add("sumto1", float(np.sum(dc * dc) + np.sum(dr * dr)), 2.0 * dc[None, :] + 2.0 * dr[:, None])
    grad = P2.T @ (G * a1[None, :]) @ P1
    return E, grad, parts


def fmap_fit_lbfgs(A, B, evals1, evals2, c00, Phi1, Phi2, a1, w_descr, w_lap, dense_weights, maxiter=5000):
    """``FunctionalMapping.fit`` with dense-map terms: scipy L-BFGS-B from x0 = c00 * e_00 with the gradient of
    column 0 zeroed (functional.py:441,477; base_functions.py:759), float64 energy (the reference evaluates it in
    float32)."""
    import scipy.optimize
    A, B = np.asarray(A, np.float64), np.asarray(B, np.float64)
    k1, k2 = A.shape[0], B.shape[0]
    Delta = ev_sqdiff(np.asarray(evals1)[:k1], np.asarray(evals2)[:k2])

    def fun(x):
        C = x.reshape(k2, k1)
        R = C @ A - B
        e = 0.5 * w_descr * np.sum(R * R) + 0.5 * w_lap * np.sum(C * C * Delta)
        g = w_descr * (R @ A.T) + w_lap * (C * Delta)
        ed, gd, _ = dense_map_energy(C, Phi1, Phi2, a1, dense_weights)
        g = g + gd
        g[:, 0] = 0.0
        return e + ed, g.ravel()

    x0 = np.zeros((k2, k1))
    x0[0, 0] = c00
    res = scipy.optimize.minimize(fun, x0.ravel(), jac=True, method="L-BFGS-B", options={"maxiter": maxiter})
    return res.x.reshape(k2, k1), res


# --------------------------------------------------------------------------------------
# the driver, on arrays
# --------------------------------------------------------------------------------------
def surface_map_arrays(Phi1, evals1, a1, Phi2, evals2, a2, c1, c2, n_ev, w_descr, w_lap,
                       icp_nit=10, nn="brute"):
    """Array-level restatement of ``compute_surface_map`` with descr+lap energy.

    densematcher/functional_map.py:44-78: preprocess (basis truncated to ``n_ev``,
    functional.py:294-295 + trimesh.py:520-523), fit (closed form, see
    ``fmap_solve_closed_form``), ``get_p2p`` (kd-tree pair, exposed as
    ``*_adjoint``, :48), dense-argmax override (:49-50), ``icp_refine`` (:71,
    nit=10), second ``get_p2p`` + override (:75-77).  The Hungarian assignment
    (:57,:78) and the precise map (:62) are outside the hot path.
    """
    k = int(n_ev)
    P1, P2 = np.asarray(Phi1, np.float64)[:, :k], np.asarray(Phi2, np.float64)[:, :k]
    l1, l2 = np.asarray(evals1, np.float64)[:k], np.asarray(evals2, np.float64)[:k]
    A = project(P1, a1, c1)
    B = project(P2, a2, c2)
    C = fmap_solve_closed_form(A, B, l1, l2, fmap_c00(P1, P2, a1, a2), w_descr, w_lap)
    out = {"A": A, "B": B, "C": C}
    p21_adj, p12_adj, MI = fm_to_p2p(C, P1, P2, a1, nn=nn)
    out["p2p_21_adjoint"], out["p2p_12_adjoint"] = p21_adj, p12_adj
    out["p2p_21"], out["p2p_12"] = dense_argmax_override(MI)
    C_icp = icp_refine(C, P1, P2, nit=icp_nit, nn=nn)
    out["C_icp"] = C_icp
    p21_adj, p12_adj, MI = fm_to_p2p(C_icp, P1, P2, a1, nn=nn)
    out["p2p_21_icp_adjoint"], out["p2p_12_icp_adjoint"] = p21_adj, p12_adj
    out["p2p_21_icp"], out["p2p_12_icp"] = dense_argmax_override(MI)
    return out


# --------------------------------------------------------------------------------------
# Hungarian assignment and the barycentric "precise map"
# --------------------------------------------------------------------------------------
def hungarian(MI, eta=None):
    """``linear_sum_assignment(MI * eta - 1000 (1 - eta), maximize=True)``, densematcher/functional_map.py:57,66,78.
    Same third-party call as the reference (scipy.optimize, 1.18.1 here)."""
    from scipy.optimize import linear_sum_assignment
    MI = np.asarray(MI, np.float64)
    eta = np.ones(MI.shape[0]) if eta is None else np.asarray(eta, np.float64)
    return linear_sum_assignment(MI * eta[..., None] - 1000 * (1 - eta[..., None]), maximize=True)



def lap_shortest_augmenting_path(cost, maximize=False):
    """Restatement of the solver behind ``scipy.optimize.linear_sum_assignment`` (third-party, scipy 1.18.1 here: the
    rectangular shortest-augmenting-path algorithm of Crouse 2016), including the details that make its answer unique
    on tied matrices and that ``dm_lap_solve`` (csrc/lap.cu) follows step for step:
      * tall matrices are transposed, maximisation negates the costs;
      * rows are added in order; each Dijkstra search scans the list ``remaining`` -- initialised to nc-1, ..., 0 and
        updated by moving its last entry into the slot of the column just scanned -- and among equal shortest-path
        costs keeps the FIRST one met, except that an unassigned column replaces an equal candidate;
      * ``r = minVal + cost[i, j] - u[i] - v[j]`` is evaluated left to right; the duals are updated as
        ``u[i] += minVal - spc[col4row[i]]``, ``v[j] -= minVal - spc[j]``.
    Returns (row_ind, col_ind, number of Dijkstra steps).  Pure Python/numpy: small matrices only."""
    cost = np.asarray(cost, np.float64)
    nr, nc = cost.shape
    transpose = nc < nr
    if transpose:
        cost = cost.T.copy()
        nr, nc = nc, nr
    if maximize:
        cost = -cost
    u, v = np.zeros(nr), np.zeros(nc)
    path = np.full(nc, -1)
    col4row, row4col = np.full(nr, -1), np.full(nc, -1)
    steps = 0
    for cur in range(nr):
        remaining = np.arange(nc - 1, -1, -1)
        nrem = nc
        SR, SC = np.zeros(nr, bool), np.zeros(nc, bool)
        spc = np.full(nc, np.inf)
        minVal, i, sink = 0.0, cur, -1
        while sink == -1:
            steps += 1
            SR[i] = True
            js = remaining[:nrem]
            r = ((minVal + cost[i, js]) - u[i]) - v[js]
            upd = r < spc[js]
            path[js[upd]] = i
            spc[js[upd]] = r[upd]
            s = spc[js]
            lowest = s.min()
            if lowest == np.inf:
                raise ValueError("cost matrix is infeasible")
            ties = np.nonzero(s == lowest)[0]
            free = ties[row4col[js[ties]] == -1]
            index = free[-1] if len(free) else ties[0]
            minVal = lowest
            j = js[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            nrem -= 1
            remaining[index] = remaining[nrem]
        u[cur] += minVal
        others = SR.copy()
        others[cur] = False
        idx = np.nonzero(others)[0]
        u[idx] += minVal - spc[col4row[idx]]
        jj = np.nonzero(SC)[0]
        v[jj] -= minVal - spc[jj]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transpose:
        order = np.argsort(col4row, kind="stable")
        return col4row[order], order, steps
    return np.arange(nr), col4row, steps

def tri_closest_point(a, b, c, d, e, f, many=False):
    """Closest point of a triangle B + s E0 + t E1 to a point P, from the six inner products
    a = E0.E0, b = E0.E1, c = E1.E1, d = E0.(B-P), e = E1.(B-P), f = (B-P).(B-P): the seven-region case analysis of
    Eberly's "Distance between point and triangle" as written in ``pointTriangleDistance``
    (pyFM/spectral/projection_utils.py:820-976).  Returns (squared distance, s, t).

    ``many=True`` reproduces the vectorised ``point_to_triangles_projection`` (:483-757), which is what the reference
    runs whenever a point has more than one candidate triangle.  It differs from the scalar routine in two region-4
    branches, where the squared distance is formed with the UN-normalised s / t of the region test (:543-544
    ``d * s + f`` and :563-564 ``e * t + f``) -- the barycentric coordinates are the same, but that (wrong) distance
    takes part in the argmin over the candidates, so it is restated as is."""
    det = a * c - b * b
    s = b * e - c * d
    t = b * d - a * e

    def full(s, t):
        return s * (a * s + b * t + 2.0 * d) + t * (b * s + c * t + 2.0 * e) + f

    if s + t <= det:
        if s < 0:
            if t < 0:                                   # region 4
                if d < 0:
                    if -d >= a:
                        return a + 2.0 * d + f, 1.0, 0.0
                    return (d * s + f if many else d * (-d / a) + f), -d / a, 0.0
                if e >= 0:
                    return f, 0.0, 0.0
                if -e >= c:
                    return c + 2.0 * e + f, 0.0, 1.0
                return (e * t + f if many else e * (-e / c) + f), 0.0, -e / c
            if e >= 0:                                  # region 3
                return f, 0.0, 0.0
            if -e >= c:
                return c + 2.0 * e + f, 0.0, 1.0
            return e * (-e / c) + f, 0.0, -e / c
        if t < 0:                                       # region 5
            if d >= 0:
                return f, 0.0, 0.0
            if -d >= a:
                return a + 2.0 * d + f, 1.0, 0.0
            return d * (-d / a) + f, -d / a, 0.0
        inv = 1.0 / det                                 # region 0
        s, t = s * inv, t * inv
        return full(s, t), s, t
    if s < 0:                                           # region 2
        tmp0, tmp1 = b + d, c + e
        if tmp1 > tmp0:
            numer, denom = tmp1 - tmp0, a - 2.0 * b + c
            if numer >= denom:
                return a + 2.0 * d + f, 1.0, 0.0
            s = numer / denom
            return full(s, 1 - s), s, 1 - s
        if tmp1 <= 0:
            return c + 2.0 * e + f, 0.0, 1.0
        if e >= 0:
            return f, 0.0, 0.0
        return e * (-e / c) + f, 0.0, -e / c
    if t < 0:                                           # region 6
        tmp0, tmp1 = b + e, a + d
        if tmp1 > tmp0:
            numer, denom = tmp1 - tmp0, a - 2.0 * b + c
            if numer >= denom:
                return c + 2.0 * e + f, 0.0, 1.0
            t = numer / denom
            return full(1 - t, t), 1 - t, t
        if tmp1 <= 0:
            return a + 2.0 * d + f, 1.0, 0.0
        if d >= 0:
            return f, 0.0, 0.0
        return d * (-d / a) + f, -d / a, 0.0
    numer = c + e - b - d                               # region 1
    if numer <= 0:
        return c + 2.0 * e + f, 0.0, 1.0
    denom = a - 2.0 * b + c
    if numer >= denom:
        return a + 2.0 * d + f, 1.0, 0.0
    s = numer / denom
    return full(s, 1 - s), s, 1 - s


def project_points_to_triangles(vert_emb, faces, points_emb, nn="kdtree"):
    """For every point the triangle of the p-dimensional mesh (vert_emb, faces) it projects onto and the barycentric
    coordinates of the projection: ``project_pc_to_triangles`` with ``precompute_dmin=True``
    (pyFM/spectral/projection_utils.py:16-115).  Per point (``project_to_mesh`` :329-377): candidate faces are those
    with  delta_min - l_max < Delta_min  (delta_min: distance to the nearest of the face's three vertices through
    the |x|^2 - 2 x.y + |y|^2 expansion, :294-326 / :191-238; l_max: longest edge, :118-146; Delta_min: distance to
    the nearest vertex, :149-186), each candidate is projected (``tri_closest_point``) and the first minimum of the
    distances wins.  Returns (face_match (n2,) int64, bary (n2, 3) float64)."""
    X, Y, faces = np.asarray(vert_emb, np.float64), np.asarray(points_emb, np.float64), np.asarray(faces)
    e0, e1, e2 = X[faces[:, 0]], X[faces[:, 1]], X[faces[:, 2]]
    lmax = np.maximum(np.maximum(np.linalg.norm(e1 - e0, axis=1), np.linalg.norm(e2 - e1, axis=1)),
                      np.linalg.norm(e0 - e2, axis=1))
    if nn == "kdtree":
        Deltamin = knn_query(X, Y, return_distance=True)[0].reshape(-1)
    else:
        Deltamin = np.linalg.norm(Y - X[knn_bruteforce(X, Y)], axis=1)
    sqX, sqY = np.linalg.norm(X, axis=1) ** 2, np.linalg.norm(Y, axis=1) ** 2
    face_match = np.zeros(Y.shape[0], dtype=np.int64)
    bary = np.zeros((Y.shape[0], 3))
    for i in range(Y.shape[0]):
        d2 = X @ Y[i]
        d2 *= -2
        d2 += sqX
        d2 += sqY[i]
        np.maximum(d2, 0, out=d2)
        dmin = np.sqrt(np.minimum(np.minimum(d2[faces[:, 0]], d2[faces[:, 1]]), d2[faces[:, 2]]))
        cand = np.nonzero(dmin - lmax < Deltamin[i])[0]
        best = (np.inf, -1, 0.0, 0.0)
        for fi in cand:
            B = X[faces[fi, 0]]
            E0, E1, D = X[faces[fi, 1]] - B, X[faces[fi, 2]] - B, B - Y[i]
            sq, s, t = tri_closest_point(E0 @ E0, E0 @ E1, E1 @ E1, E0 @ D, E1 @ D, D @ D, many=len(cand) > 1)
            dist = np.sqrt(max(sq, 0.0))
            if dist < best[0]:
                best = (dist, fi, s, t)
        face_match[i] = best[1]
        bary[i] = (1 - best[2] - best[3], best[2], best[3])
    return face_match, bary


def fm_to_precise_map(C, Phi1, Phi2, faces1, use_adj=True, nn="kdtree"):
    """``mesh_FM_to_p2p_precise`` (pyFM/spectral/convert.py:186-231) + ``barycentric_to_precise``
    (projection_utils.py:380-417): the (n2, n1) sparse map whose row i holds the barycentric coordinates of the image
    of vertex i of mesh 2 on a triangle of mesh 1.  Returns (csr matrix, face_match, bary)."""
    import scipy.sparse as sp
    C = np.asarray(C, np.float64)
    k2, k1 = C.shape
    if use_adj:
        emb1, emb2 = Phi1[:, :k1], Phi2[:, :k2] @ C
    else:
        emb1, emb2 = Phi1[:, :k1] @ C.T, Phi2[:, :k2]
    fm, bary = project_points_to_triangles(emb1, faces1, emb2, nn=nn)
    faces1 = np.asarray(faces1)
    n2, n1 = emb2.shape[0], emb1.shape[0]
    I = np.tile(np.arange(n2), 3)
    J = np.concatenate([faces1[fm, 0], faces1[fm, 1], faces1[fm, 2]])
    S = np.concatenate([bary[:, 0], bary[:, 1], bary[:, 2]])
    return sp.csr_matrix((S, (I, J)), shape=(n2, n1)), fm, bary
