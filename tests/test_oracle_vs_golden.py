"""Pin the CPU oracle (oracle/dm_oracle.py) against answers produced by the reference itself.

The fixtures in tests/golden/ were minted by oracle/make_goldens.py from the
unmodified reference (knn_query, FM_to_p2p, p2p_to_FM, icp_refine,
compute_surface_map); arrays named ``ref_*`` are reference outputs.
"""
import numpy as np
import pytest

from oracle import dm_oracle as orc, meshgen


def relF(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_cfg1_inputs_regenerate(golden_nn_cfg1):
    g = golden_nn_cfg1
    F1 = meshgen.random_unit_features(int(g["n1"]), int(g["d"]), np.random.default_rng(int(g["seed1"])))
    F2 = meshgen.random_unit_features(int(g["n2"]), int(g["d"]), np.random.default_rng(int(g["seed2"])))
    assert F1.astype(np.float64).sum() == g["checksum1"]
    assert F2.astype(np.float64).sum() == g["checksum2"]


def test_nn_cfg1_cosine_argmax_equals_reference_kdtree(golden_nn_cfg1):
    g = golden_nn_cfg1
    F1 = meshgen.random_unit_features(2000, 384, np.random.default_rng(int(g["seed1"])))
    F2 = meshgen.random_unit_features(2000, 384, np.random.default_rng(int(g["seed2"])))
    # Euclidean form (exact restatement of knn_query) ...
    assert np.array_equal(orc.knn_bruteforce(F1, F2), g["ref_p2p_21"])
    assert np.array_equal(orc.knn_bruteforce(F2, F1), g["ref_p2p_12"])
    # ... and the cosine form the north star names (rows are unit norm, model.py:169)
    assert np.array_equal(orc.nn_argmax(F2, F1), g["ref_p2p_21"])
    # column-argmax of the same score matrix gives the reverse map on unit rows
    assert np.array_equal(orc.nn_argmax(F2, F1, axis=0), g["ref_p2p_12"])


def test_nn_small_nonunit_rows_and_distances(golden_nn_small):
    g = golden_nn_small
    X, Y = g["X"], g["Y"]
    assert np.array_equal(orc.knn_bruteforce(X, Y), g["ref_match"])
    d, m = orc.knn_query(X, Y, return_distance=True)
    assert np.array_equal(m, g["ref_match"]) and np.allclose(d, g["ref_dist"], rtol=0, atol=1e-12)
    d3, m3 = orc.knn_query(X, Y, k=3, return_distance=True)
    assert np.array_equal(m3, g["ref_match_k3"]) and m3.shape == (193, 3)
    best = np.linalg.norm(Y.astype(np.float64) - X.astype(np.float64)[m], axis=1)
    assert np.allclose(best, g["ref_dist"], atol=1e-12)


def test_nn_duplicate_rows_lowest_index():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((64, 16)).astype(np.float32)
    X[40] = X[7]
    X[41] = X[7]
    Y = X[[7, 40, 41, 3]]
    assert orc.nn_argmax(Y, X).tolist() == [7, 7, 7, 3]
    assert orc.knn_bruteforce(X, Y).tolist() == [7, 7, 7, 3]


def test_closed_form_matches_reference_lbfgs(golden_fm):
    g = golden_fm
    k = int(g["k"])
    A = orc.project(g["Phi1"], g["area1"], g["c1"], k)
    B = orc.project(g["Phi2"], g["area2"], g["c2"], k)
    c00 = orc.fmap_c00(g["Phi1"], g["Phi2"], g["area1"], g["area2"])
    C = orc.fmap_solve_closed_form(A, B, g["evals1"], g["evals2"], c00, float(g["w_descr"]), float(g["w_lap"]))
    assert relF(C, g["C_closed_form"]) < 1e-12
    # the reference's L-BFGS-B run (fp32 energy) agrees to its own optimiser noise
    assert relF(g["ref_C_lbfgs"], C) < 1e-3
    assert C[0, 0] == pytest.approx(c00) and np.all(C[1:, 0] == 0)
    # stationarity: gradient of the free block vanishes
    Delta = orc.ev_sqdiff(g["evals1"][:k], g["evals2"][:k])
    grad = float(g["w_descr"]) * (C @ A - B) @ A.T + float(g["w_lap"]) * C * Delta
    assert np.abs(grad[:, 1:]).max() < 1e-6 * np.abs(float(g["w_descr"]) * B @ A.T).max()
    e0 = orc.fmap_energy(C, A, B, Delta, float(g["w_descr"]), float(g["w_lap"]))
    assert e0 <= orc.fmap_energy(g["ref_C_lbfgs"], A, B, Delta, float(g["w_descr"]), float(g["w_lap"])) * (1 + 1e-6)


@pytest.mark.parametrize("nn", ["brute", "tree"])
def test_fm_to_p2p_matches_reference(golden_fm, nn):
    g = golden_fm
    p21, p12, MI = orc.fm_to_p2p(g["C_closed_form"], g["Phi1"], g["Phi2"], g["area1"], nn=nn)
    assert np.array_equal(p21, g["ref_cf_p2p_21"])
    assert np.array_equal(p12, g["ref_cf_p2p_12"])
    assert np.allclose(MI[:8, :8], g["ref_cf_MI_corner"], rtol=1e-12, atol=1e-14)
    assert MI.sum() == pytest.approx(float(g["ref_cf_MI_sum"]), rel=1e-10)
    assert np.linalg.norm(MI) == pytest.approx(float(g["ref_cf_MI_fro"]), rel=1e-12)
    a1, a0 = orc.dense_argmax_override(MI, np.ones(MI.shape[0]))
    assert np.array_equal(a1, g["ref_cf_MI_argmax1"]) and np.array_equal(a0, g["ref_cf_MI_argmax0"])


def test_dense_argmax_is_a_scaled_score_argmax(golden_fm):
    """the dense override equals argmax_j S_ij a1[j] / argmax_i S_ij."""
    g = golden_fm
    C, P1, P2, a1 = g["C_closed_form"], g["Phi1"], g["Phi2"], g["area1"]
    assert np.array_equal(orc.nn_argmax(P2 @ C, P1, col_scale=a1), g["ref_cf_MI_argmax1"])
    assert np.array_equal(orc.nn_argmax(P2 @ C, P1, axis=0), g["ref_cf_MI_argmax0"])


def test_p2p_to_fm_matches_reference(golden_fm):
    g = golden_fm
    p = g["ref_cf_p2p_21"]
    assert relF(orc.p2p_to_fm(p, g["Phi1"], g["Phi2"], A2=g["area2"]), g["ref_cf_C_area"]) < 1e-13
    import scipy.sparse as sp
    assert relF(orc.p2p_to_fm(p, g["Phi1"], g["Phi2"], A2=sp.diags(g["area2"]).tocsc()), g["ref_cf_C_area"]) < 1e-13
    assert relF(orc.p2p_to_fm(p, g["Phi1"], g["Phi2"]), g["ref_cf_C_lstsq"]) < 1e-12
    with pytest.raises(ValueError):
        orc.p2p_to_fm(p, g["Phi1"], g["Phi2"], A2=g["area2"][:-1])


def test_icp_matches_reference(golden_fm):
    g = golden_fm
    C, p = orc.icp_refine(g["C_closed_form"], g["Phi1"], g["Phi2"], nit=10, return_p2p=True)
    assert relF(C, g["ref_cf_C_icp"]) < 1e-10
    assert np.array_equal(p, g["ref_cf_p2p_icp"])
    assert np.allclose(C @ C.T, np.eye(C.shape[0]), atol=1e-10)


def test_surface_map_arrays_vs_reference_driver(golden_fm):
    """End to end vs compute_surface_map.  C differs by the reference's L-BFGS noise, so index
    maps may differ on a handful of vertices; the bound here is 1 %."""
    g = golden_fm
    out = orc.surface_map_arrays(g["Phi1"], g["evals1"], g["area1"], g["Phi2"], g["evals2"], g["area2"],
                                 g["c1"], g["c2"], int(g["k"]), float(g["w_descr"]), float(g["w_lap"]))
    assert relF(out["C"], g["ref_C_lbfgs"]) < 1e-3
    n = len(g["ref_p2p_21"])
    for key in ("p2p_21", "p2p_12", "p2p_21_adjoint", "p2p_12_adjoint"):
        assert np.count_nonzero(out[key] != g["ref_" + key]) <= 0.01 * n, key
    for key in ("p2p_21_icp", "p2p_12_icp", "p2p_21_icp_adjoint", "p2p_12_icp_adjoint"):
        assert np.count_nonzero(out[key] != g["ref_" + key]) <= 0.05 * n, key


def test_zoomout_upstream_semantics(golden_zo):
    g = golden_zo
    assert bool(g["ref_shipped_zoomout_raises_typeerror"])
    for nn in ("brute", "tree"):
        C, p = orc.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=14, step=1, A2=g["area2"],
                                  return_p2p=True, nn=nn)
        assert C.shape == (26, 26) and relF(C, g["ref_C_zo"]) < 1e-12
        assert np.array_equal(p, g["ref_p2p_zo"])
    C, p = orc.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=9, step=(2, 3), A2=g["area2"], return_p2p=True)
    assert C.shape == (39, 30) and relF(C, g["ref_C_zo_rect"]) < 1e-12
    assert np.array_equal(p, g["ref_p2p_zo_rect"])
    C = orc.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=6, step=1, A2=g["area2"],
                           subsample=(g["sub1"], g["sub2"]))
    assert relF(C, g["ref_C_zo_sub"]) < 1e-10
    with pytest.raises(AssertionError):
        orc.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=40, step=1, A2=g["area2"])


def test_meshgen_basis_is_area_orthonormal():
    V, F = meshgen.icosphere(2)
    ev, Phi, a = meshgen.lbo_basis(meshgen.deform(V, (1, 1.2, 0.8)), F, 25)
    assert np.allclose(Phi.T @ (a[:, None] * Phi), np.eye(25), atol=1e-8)
    assert abs(ev[0]) < 1e-6 and np.all(np.diff(ev) > -1e-9)
    ev, Phi, a = meshgen.synthetic_basis(500, 30, np.random.default_rng(0))
    assert np.allclose(Phi.T @ (a[:, None] * Phi), np.eye(30), atol=1e-9)
    assert np.allclose(Phi[:, 0], Phi[0, 0])


def test_dense_map_energy_terms_match_reference_torch():
    """The reference's own torch implementations of the dense-map terms (base_functions.py:296-428), evaluated in
    float64 at a fixed C in the authoring container, pin the oracle's values and analytic gradients; the oracle's
    L-BFGS fit with the notebook's weights lands on the reference's fit (within the reference's float32 noise)."""
    from conftest import load_golden
    from oracle import dm_oracle as orc
    g, e = load_golden("fm_pair_ico3.npz"), load_golden("energy_ico3.npz")
    k = int(e["k"])
    P1, P2, a1, a2 = g["Phi1"][:, :k], g["Phi2"][:, :k], g["area1"], g["area2"]
    for name in orc.DENSE_TERMS:
        E, G, parts = orc.dense_map_energy(e["C"], P1, P2, a1, {name: 1.0})
        assert np.isclose(parts[name], float(e["ref_E_" + name]), rtol=1e-12), name
        assert np.abs(G - e["ref_G_" + name]).max() < 1e-10 * np.abs(e["ref_G_" + name]).max(), name
    A, B = orc.project(P1, a1, g["c1"]), orc.project(P2, a2, g["c2"])
    w = dict(ent=float(e["w_ent"]), sumto1=float(e["w_sumto1"]))
    C, res = orc.fmap_fit_lbfgs(A, B, g["evals1"], g["evals2"], orc.fmap_c00(P1, P2, a1, a2), P1, P2, a1,
                                float(e["w_descr"]), float(e["w_lap"]), w)
    Cr = e["ref_C_notebook"]
    assert Cr.shape == (k, k) and np.linalg.norm(C - Cr) / np.linalg.norm(Cr) < 1e-3
    assert np.array_equal(orc.fm_to_p2p(C, P1, P2, a1)[0], orc.fm_to_p2p(Cr, P1, P2, a1)[0])


def _csr(g, prefix, shape):
    import scipy.sparse as sp
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=shape)


def test_precise_map_matches_reference(golden_fm, golden_extras):
    """8f rank 2: the oracle's barycentric precise map against the reference's project_pc_to_triangles."""
    g, x = golden_fm, golden_extras
    k = int(g["k"])
    P, fm_, bary = orc.fm_to_precise_map(g["C_closed_form"], g["Phi1"][:, :k], g["Phi2"][:, :k], x["faces"])
    ref = _csr(x, "ref_precise", (642, 642))
    assert abs(P - ref).max() < 1e-12
    assert np.allclose(bary.sum(1), 1.0) and bary.min() > -1e-12
    # every region of the case analysis incl. the two branches where the vectorised routine differs
    Pr = orc.project_points_to_triangles(x["rnd_X"], x["rnd_F"], x["rnd_Y"])
    F = x["rnd_F"]
    import scipy.sparse as sp
    n = len(x["rnd_Y"])
    got = sp.csr_matrix((Pr[1].T.ravel(), (np.tile(np.arange(n), 3), F[Pr[0]].T.ravel())), shape=(n, 40))
    assert abs(got - _csr(x, "ref_rnd", (n, 40))).max() < 1e-12
    # brute-force nearest vertex instead of the kd-tree gives the same map
    Pb = orc.project_points_to_triangles(x["rnd_X"], x["rnd_F"], x["rnd_Y"], nn="brute")
    assert np.array_equal(Pb[0], Pr[0]) and np.array_equal(Pb[1], Pr[1])


def test_hungarian_matches_reference(golden_fm, golden_extras):
    g, x = golden_fm, golden_extras
    k = int(g["k"])
    _, _, MI = orc.fm_to_p2p(g["C_closed_form"], g["Phi1"][:, :k], g["Phi2"][:, :k], g["area1"])
    r, c = orc.hungarian(MI)
    assert np.array_equal(r, x["ref_hungarian_rows"])
    # the indicator is reproduced to rounding, so the optimal assignment (unique here) must agree
    assert np.array_equal(c, x["ref_hungarian_cols"])
    rp, cp = orc.hungarian(_csr(x, "ref_precise", (642, 642)).toarray())
    assert np.array_equal(cp, x["ref_hungarian_precise_cols"])


def test_lap_restatement_equals_scipy_including_ties():
    """The solver dm_lap_solve follows (oracle restatement) against scipy itself: generic, integer-tie, constant and
    rectangular matrices, both directions of optimisation."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(0)
    cases = [rng.standard_normal((50, 50)), rng.standard_normal((40, 60)), rng.standard_normal((60, 40)),
             rng.integers(0, 4, (30, 30)).astype(float), rng.integers(0, 4, (30, 45)).astype(float),
             rng.integers(0, 4, (45, 30)).astype(float), rng.integers(0, 2, (64, 64)).astype(float), np.ones((20, 20)),
             np.zeros((9, 17))]
    for c in cases:
        for mx in (False, True):
            r, col, _ = orc.lap_shortest_augmenting_path(c, mx)
            rr, cc = linear_sum_assignment(c, maximize=mx)
            assert np.array_equal(r, rr) and np.array_equal(col, cc)
    inf = np.full((4, 4), np.inf)
    inf[:, 0] = 1.0
    with pytest.raises(ValueError):
        orc.lap_shortest_augmenting_path(inf)


# ---------------------------------------------------------------------------------------------- BASELINE size (N = 2562, k = 100)
def _zo_hash(p):
    return int((p.astype(np.int64) * (np.arange(len(p)) % 1009 + 1)).sum())


def test_full_size_fm_to_p2p_and_p2p_to_fm_equal_the_reference(golden_full):
    g = golden_full
    k = int(g["k"])
    P1, P2 = g["Phi1"][:, :k], g["Phi2"][:, :k]
    C = g["C_closed_form"]
    # the closed-form C stored in the file is what the oracle computes from the stored inputs
    A, B = orc.project(P1, g["area1"], g["c1"]), orc.project(P2, g["area2"], g["c2"])
    Co = orc.fmap_solve_closed_form(A, B, g["evals1"][:k], g["evals2"][:k], orc.fmap_c00(P1, P2, g["area1"], g["area2"]),
                                    float(g["w_descr"]), float(g["w_lap"]))
    assert relF(Co, C) < 1e-12
    p21, p12, MI = orc.fm_to_p2p(C, P1, P2, g["area1"])
    assert np.array_equal(p21, g["ref_p2p_21"]) and np.array_equal(p12, g["ref_p2p_12"])
    assert np.array_equal(MI.argmax(1), g["ref_MI_argmax1"]) and np.array_equal(MI.argmax(0), g["ref_MI_argmax0"])
    assert abs(MI.sum() - g["ref_MI_sum"]) < 1e-8 * abs(g["ref_MI_sum"]) + 1e-8
    assert relF(orc.p2p_to_fm(g["ref_p2p_21"], P1, P2, g["area2"]), g["ref_C_area"]) < 1e-13


def test_full_size_icp_equals_the_reference(golden_full):
    g = golden_full
    k = int(g["k"])
    C, p = orc.icp_refine(g["C_closed_form"], g["Phi1"][:, :k], g["Phi2"][:, :k], nit=10, return_p2p=True)
    assert relF(C, g["ref_C_icp"]) < 1e-9
    assert np.array_equal(p, g["ref_p2p_icp"])


def test_full_size_zoomout_ladder_equals_the_reference_primitives(golden_full):
    """30 -> 200, 170 rungs: the brute-force float64 oracle walks the same p2p sequence as the ladder composed from the
    reference's kd-tree knn_query + p2p_to_FM, rung by rung (hashes) and at the end (C, p2p)."""
    g = golden_full
    C = g["C_closed_form"][:30, :30].copy()
    P1, P2, a2 = g["Phi1"], g["Phi2"], g["area2"]
    for it in range(170):
        kk = C.shape[0]
        p = orc.knn_bruteforce(P1[:, :kk] @ C.T, P2[:, :kk])   # upstream p2p_21: knn(tree = Phi1 C^T, query = Phi2)
        assert _zo_hash(p) == int(g["ref_zo_p2p_hashes"][it]), it
        C = orc.p2p_to_fm(p, P1[:, :kk + 1], P2[:, :kk + 1], a2)
    assert relF(C, g["ref_C_zo"]) < 1e-12
    assert np.array_equal(orc.knn_bruteforce(P1 @ C.T, P2), g["ref_p2p_zo"])
