"""GPU barycentric precise map (dm_precise_map) against the reference-minted golden and the oracle, and the
compute_extra slots of compute_surface_map (Hungarian on the mapped indicator and on the precise map)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import dm_oracle as orc

pytestmark = pytest.mark.gpu


def _csr(g, prefix, shape):
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], g[prefix + "_indptr"]), shape=shape)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def test_random_mesh_all_regions_match_reference(golden_extras):
    from densematcher_b200.pyFM.spectral import projection_utils as pju
    x = golden_extras
    P, face, bary = pju.project_pc_to_triangles(x["rnd_X"], x["rnd_F"], x["rnd_Y"], return_bary=True)
    n = len(x["rnd_Y"])
    assert abs(P - _csr(x, "ref_rnd", (n, 40))).max() < 1e-12
    # Face ids are compared through the map they define: a point whose projection is a mesh vertex or lies on an edge
    # is equidistant (to rounding) from every face sharing it (and this random soup repeats triangles), and which of
    # them the argmin reports is decided by the last bit of six inner products -- in the reference as much as here.
    fo, bo = orc.project_points_to_triangles(x["rnd_X"], x["rnd_F"], x["rnd_Y"])
    assert abs(P - pju.barycentric_to_precise(x["rnd_F"], fo, bo, 40)).max() < 1e-12
    assert np.abs(bary.sum(1) - 1).max() < 1e-12
    assert face.dtype == np.int64 and sp.isspmatrix_csr(P)


def test_spectral_precise_map_matches_reference(golden_fm, golden_extras):
    from densematcher_b200.pyFM import spectral
    from densematcher_b200.pyFM.mesh import TriMesh
    g, x = golden_fm, golden_extras
    k = int(g["k"])
    m1 = TriMesh.from_basis(g["evals1"], g["Phi1"], g["area1"], faces=x["faces"])
    m2 = TriMesh.from_basis(g["evals2"], g["Phi2"], g["area2"], faces=x["faces"])
    P = spectral.mesh_FM_to_p2p_precise(g["C_closed_form"], m1, m2)
    ref = _csr(x, "ref_precise", (642, 642))
    assert P.shape == (642, 642) and abs(P - ref).max() < 1e-11
    # on a manifold mesh, points projecting strictly inside a triangle have a well-defined face id
    _, face, bary = spectral.projection_utils.project_pc_to_triangles(g["Phi1"][:, :k], x["faces"],
                                                                      g["Phi2"][:, :k] @ g["C_closed_form"], return_bary=True)
    _, fo, bo = orc.fm_to_precise_map(g["C_closed_form"], g["Phi1"][:, :k], g["Phi2"][:, :k], x["faces"])
    interior = bo.min(1) > 1e-9
    assert interior.sum() > 100 and np.array_equal(face[interior], fo[interior])
    assert np.abs(bary[interior] - bo[interior]).max() < 1e-11
    # the other branch (use_adj=False) against the oracle
    P2 = spectral.mesh_FM_to_p2p_precise(g["C_closed_form"], m1, m2, use_adj=False)
    Po, _, _ = orc.fm_to_precise_map(g["C_closed_form"], g["Phi1"][:, :k], g["Phi2"][:, :k], x["faces"], use_adj=False)
    assert abs(P2 - Po).max() < 1e-11


def test_ragged_batch_equals_single_calls(golden_extras):
    from densematcher_b200 import fm
    x = golden_extras
    X, F, Y = x["rnd_X"], x["rnd_F"], x["rnd_Y"]
    rng = np.random.default_rng(5)
    X2 = rng.standard_normal((33, 5))
    F2 = np.stack([rng.choice(33, 3, replace=False) for _ in range(50)]).astype(np.int32)
    Y2 = 1.5 * rng.standard_normal((77, 5))
    face, bary = fm.precise_map(dev(np.concatenate([X, X2])), dev(np.concatenate([F, F2])), dev(np.concatenate([Y, Y2])),
                                off1=[0, 40, 73], face_off=[0, len(F), len(F) + 50], off2=[0, len(Y), len(Y) + 77],
                                out_dtype=torch.int32)
    face, bary = face.cpu().numpy(), bary.cpu().numpy()
    assert face.dtype == np.int32
    from densematcher_b200.pyFM.spectral.projection_utils import barycentric_to_precise as b2p
    for (Xi, Fi, Yi, sl) in ((X, F, Y, slice(0, len(Y))), (X2, F2, Y2, slice(len(Y), len(Y) + 77))):
        fo, bo = orc.project_points_to_triangles(Xi, Fi, Yi, nn="brute")
        assert abs(b2p(Fi, face[sl], bary[sl], len(Xi)) - b2p(Fi, fo, bo, len(Xi))).max() < 1e-12


def test_compute_extra_slots_match_reference(golden_fm, golden_extras):
    """compute_surface_map(compute_extra=True): Hungarian on the mapped indicator and on the precise map
    (functional_map.py:57-66), both solved on the GPU, against scipy run on the reference's own matrices."""
    from densematcher_b200 import _lib
    from densematcher_b200.functional_map import compute_surface_map
    from densematcher_b200.pyFM import FunctionalMapping
    from densematcher_b200.pyFM.mesh import TriMesh
    g, x = golden_fm, golden_extras
    k = int(g["k"])
    meshes = [TriMesh.from_basis(g["evals1"], g["Phi1"], g["area1"], faces=x["faces"]),
              TriMesh.from_basis(g["evals2"], g["Phi2"], g["area2"], faces=x["faces"])]
    FunctionalMapping.projection_flags = _lib.DM_F64_GEMM
    try:
        out = compute_surface_map(meshes[0], meshes[1], g["c1"], g["c2"], n_ev=k, compute_extra=True,
                                  fit_params=dict(w_descr=float(g["w_descr"]), w_lap=float(g["w_lap"]), w_dcomm=0))
    finally:
        FunctionalMapping.projection_flags = 0
    assert np.array_equal(out[2][0], x["ref_hungarian_rows"]) and np.array_equal(out[2][1], x["ref_hungarian_cols"])
    assert np.array_equal(out[3][0], x["ref_hungarian_precise_rows"])
    assert np.array_equal(out[3][1], x["ref_hungarian_precise_cols"])
    # hungarian_icp: identical to scipy on the materialised indicator
    from scipy.optimize import linear_sum_assignment
    r, c = linear_sum_assignment(out[7].mapped_indicator, maximize=True)
    assert np.array_equal(out[6][0], r) and np.array_equal(out[6][1], c)


def test_batched_hungarian_and_precise_maps(golden_fm, golden_extras):
    """pipeline.hungarian_pairs / precise_maps on a two-pair batch: every pair equals the reference-minted golden."""
    from densematcher_b200 import pipeline
    from densematcher_b200.pyFM.spectral.projection_utils import barycentric_to_precise
    g, x = golden_fm, golden_extras
    k = int(g["k"])
    two = lambda a: np.concatenate([a, a])
    off = np.array([0, 642, 1284])
    batch = pipeline.PairBatchDevice(
        dev(two(g["c1"])), dev(two(g["c2"])), off, off, torch.device("cuda"),
        Phi1=dev(two(g["Phi1"][:, :k])), Phi2=dev(two(g["Phi2"][:, :k])), evals1=dev(two(g["evals1"][None, :k])),
        evals2=dev(two(g["evals2"][None, :k])), area1=dev(two(g["area1"])), area2=dev(two(g["area2"])))
    C = dev(np.stack([g["C_closed_form"]] * 2))
    for r, c in pipeline.hungarian_pairs(batch, C):
        assert np.array_equal(r, x["ref_hungarian_rows"]) and np.array_equal(c, x["ref_hungarian_cols"])
    faces = x["faces"].astype(np.int32)
    face, bary = pipeline.precise_maps(batch, C, np.concatenate([faces, faces]), [0, len(faces), 2 * len(faces)])
    face, bary = face.cpu().numpy(), bary.cpu().numpy()
    ref = _csr(x, "ref_precise", (642, 642))
    for p in range(2):
        P = barycentric_to_precise(faces, face[642 * p:642 * (p + 1)], bary[642 * p:642 * (p + 1)], 642)
        assert abs(P - ref).max() < 1e-11


def test_batched_mapped_indicators_ragged(golden_fm):
    """dm_mapped_indicators (one call for a ragged batch) gives, bit for bit, the matrices of the per-pair
    dm_mapped_indicator, and hungarian_pairs on the ragged batch the assignments of the per-pair solve."""
    from densematcher_b200 import fm, pipeline
    g = golden_fm
    k = int(g["k"])
    rng = np.random.default_rng(7)
    n1s, n2s = [642, 300, 511], [642, 420, 200]
    P1 = np.concatenate([g["Phi1"][:n, :k] for n in n1s]); P2 = np.concatenate([g["Phi2"][:n, :k] for n in n2s])
    a1 = np.concatenate([g["area1"][:n] for n in n1s])
    off1, off2 = np.concatenate([[0], np.cumsum(n1s)]), np.concatenate([[0], np.cumsum(n2s)])
    C = np.stack([g["C_closed_form"], g["C_closed_form"] + 0.01 * rng.standard_normal((k, k)), np.eye(k)])
    mats = fm.mapped_indicators(dev(C), dev(P1), dev(P2), dev(a1), off1, off2)
    singles = []
    for p in range(3):
        one = fm.mapped_indicator(dev(C[p]), dev(P1[off1[p]:off1[p + 1]]), dev(P2[off2[p]:off2[p + 1]]), dev(a1[off1[p]:off1[p + 1]]))
        assert mats[p].shape == one.shape and torch.equal(mats[p], one), p
        singles.append(one)
    batch = pipeline.PairBatchDevice(
        dev(np.zeros((int(off1[-1]), 4), np.float32)), dev(np.zeros((int(off2[-1]), 4), np.float32)), off1, off2,
        torch.device("cuda"), Phi1=dev(P1), Phi2=dev(P2), area1=dev(a1))
    want = fm.lap_solve(singles, maximize=True)
    for chunk in (None, 2):
        got = pipeline.hungarian_pairs(batch, dev(C), chunk_pairs=chunk)
        for (r, c), (rw, cw) in zip(got, want):
            assert np.array_equal(r, rw) and np.array_equal(c, cw)
