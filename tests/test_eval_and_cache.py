"""CPU tests of the small pieces around the hot path: evaluation metrics and the operator-cache loader."""
import numpy as np
import scipy.sparse as sp

from oracle import meshgen


def test_metrics_match_their_definitions():
    from densematcher_b200.pyFM.eval import accuracy, continuity, coverage
    rng = np.random.default_rng(0)
    n1, n2 = 30, 25
    X1, X2 = rng.random((n1, 3)), rng.random((n2, 3))
    D1 = np.linalg.norm(X1[:, None] - X1[None], axis=2); D2 = np.linalg.norm(X2[:, None] - X2[None], axis=2)
    p, gt = rng.integers(0, n1, n2), rng.integers(0, n1, n2)
    acc, all_d = accuracy(p, gt, D1, return_all=True, sqrt_area=2.0)
    assert np.allclose(all_d, [D1[a, b] / 2.0 for a, b in zip(p, gt)]) and np.isclose(acc, all_d.mean())
    edges = np.array([[0, 1], [1, 2], [3, 7]])
    assert np.isclose(continuity(p, D1, D2, edges), np.mean([D1[p[a], p[b]] / D2[a, b] for a, b in edges]))
    a = rng.random(n1)
    assert np.isclose(coverage(p, a), a[np.unique(p)].sum() / a.sum())
    assert np.isclose(coverage(p, sp.diags(a).tocsr()), a[np.unique(p)].sum() / a.sum())


def test_operator_cache_roundtrip(tmp_path):
    from densematcher_b200.spectral_ops import load_operator_cache
    V, F = meshgen.icosphere(1)
    evals, Phi, area = meshgen.lbo_basis(V, F, 10)
    L = sp.identity(V.shape[0], format="csr", dtype=np.float32)
    path = tmp_path / "abc_0.npz"
    np.savez(path, verts=V.astype(np.float32), faces=F, k_eig=10, frames=np.zeros((V.shape[0], 3, 3), np.float32),
             mass=area.astype(np.float32), evals=evals.astype(np.float32), evecs=Phi.astype(np.float32),
             L_data=L.data, L_indices=L.indices, L_indptr=L.indptr, L_shape=L.shape)
    m = load_operator_cache(path, k_eig=6)
    assert m.eigenvectors.shape == (V.shape[0], 6) and m.eigenvectors.dtype == np.float64
    assert np.allclose(m.eigenvalues, evals[:6], atol=1e-5) and np.allclose(m.vertex_areas, area, rtol=1e-6)
    assert m.process(4).eigenvectors.shape[1] == 4          # slicing an existing spectrum needs no GPU and no geometry


def test_precise_map_host_helpers():
    from densematcher_b200.pyFM import spectral
    assert callable(spectral.mesh_FM_to_p2p_precise) and callable(spectral.projection_utils.project_pc_to_triangles)
    P = spectral.projection_utils.barycentric_to_precise(np.array([[0, 1, 2], [1, 2, 3]]), np.array([1, 0, 1]),
                                                          np.array([[0.2, 0.3, 0.5], [1.0, 0, 0], [0, 0, 1.0]]), 5)
    assert P.shape == (3, 5) and np.allclose(P.toarray()[0], [0, 0.2, 0.3, 0.5, 0])
