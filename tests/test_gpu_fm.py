"""GPU parity of the functional-map stages against the oracle and the reference-minted goldens.
Index outputs bit-exact; C within 1e-4 relative Frobenius (north star), in practice ~1e-12."""
import numpy as np
import pytest
import torch

from oracle import dm_oracle as orc, meshgen

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope="module")
def fm():
    from densematcher_b200 import fm as _fm
    return _fm


def test_projection_and_closed_form_solve(fm, golden_fm):
    from densematcher_b200 import _lib
    g = golden_fm
    k = int(g["k"])
    Ao, Bo = orc.project(g["Phi1"], g["area1"], g["c1"], k), orc.project(g["Phi2"], g["area2"], g["c2"], k)
    c00 = orc.fmap_c00(g["Phi1"], g["Phi2"], g["area1"], g["area2"])
    # float64 contraction: agrees with numpy to rounding; tcgen05 split-bf16 contraction (default): fp32-grade,
    # like the reference's own float32 projection (base_functions.py:516-532)
    for flags, tolA, tolC in ((_lib.DM_F64_GEMM, 1e-12, 1e-9), (0, 5e-6, 1e-4)):
        A = fm.project(dev(g["Phi1"]), dev(g["area1"]), dev(g["c1"]), k=k, flags=flags)[0].cpu().numpy()
        B = fm.project(dev(g["Phi2"]), dev(g["area2"]), dev(g["c2"]), k=k, flags=flags)[0].cpu().numpy()
        assert relF(A, Ao) < tolA and relF(B, Bo) < tolA
        C = fm.fmap_solve(dev(A)[None], dev(B)[None], dev(g["evals1"][:k])[None], dev(g["evals2"][:k])[None],
                          dev(np.array([c00])), float(g["w_descr"]), float(g["w_lap"]))[0].cpu().numpy()
        assert relF(C, g["C_closed_form"]) < tolC          # bar: 1e-4 (north star)
        assert relF(C, g["ref_C_lbfgs"]) < 1e-3            # the reference's own L-BFGS noise
        assert C[0, 0] == c00 and np.all(C[1:, 0] == 0)


def test_projection_tensor_core_engine_shapes(fm):
    """tcgen05 projection on ragged batches, k above one 128-row tile, d not a multiple of 64, tiny meshes."""
    rng = np.random.default_rng(21)
    for ns, k, d in (((2000, 1963, 1500), 100, 384), ((700, 33), 130, 200), ((17,), 5, 7), ((2500, 80, 64), 64, 512)):
        bases = [meshgen.synthetic_basis(n, min(k, n), rng) for n in ns]
        kk = min(k, min(ns))
        Phi = np.concatenate([b[1][:, :kk] for b in bases]); area = np.concatenate([b[2] for b in bases])
        F = meshgen.random_unit_features(sum(ns), d, rng)
        off = np.concatenate([[0], np.cumsum(ns)])
        out = fm.project(dev(Phi), dev(area), dev(F), off, k=kk).cpu().numpy()
        for i in range(len(ns)):
            s = slice(off[i], off[i + 1])
            ref = orc.project(Phi[s], area[s], F[s], kk)
            assert np.abs(out[i] - ref).max() < 4e-6 * np.abs(ref).max() + 1e-12, (ns, k, d, i)


def test_fm_to_p2p_four_outputs(fm, golden_fm):
    g = golden_fm
    out = fm.fm_to_p2p(dev(g["C_closed_form"]), dev(g["Phi1"]), dev(g["Phi2"]), dev(g["area1"]))
    assert np.array_equal(out["p2p_21"].cpu().numpy(), g["ref_cf_p2p_21"])
    assert np.array_equal(out["p2p_12"].cpu().numpy(), g["ref_cf_p2p_12"])
    assert np.array_equal(out["dense_21"].cpu().numpy(), g["ref_cf_MI_argmax1"])
    assert np.array_equal(out["dense_12"].cpu().numpy(), g["ref_cf_MI_argmax0"])
    MI = fm.mapped_indicator(dev(g["C_closed_form"]), dev(g["Phi1"]), dev(g["Phi2"]), dev(g["area1"])).cpu().numpy()
    assert np.allclose(MI[:8, :8], g["ref_cf_MI_corner"], rtol=1e-11, atol=1e-13)
    assert np.linalg.norm(MI) == pytest.approx(float(g["ref_cf_MI_fro"]), rel=1e-12)


def test_reference_facing_FM_to_p2p_and_p2p_to_FM(golden_fm):
    import scipy.sparse as sp
    from densematcher_b200.pyFM import spectral
    g = golden_fm
    A1 = sp.diags(g["area1"]).tocsc()
    p21, p12, MI = spectral.FM_to_p2p(g["C_closed_form"], g["Phi1"], g["Phi2"], A1)
    assert p21.dtype == np.int64 and MI.shape == (642, 642)
    assert np.array_equal(p21, g["ref_cf_p2p_21"]) and np.array_equal(p12, g["ref_cf_p2p_12"])
    assert np.array_equal(MI.argmax(1), g["ref_cf_MI_argmax1"]) and np.array_equal(MI.argmax(0), g["ref_cf_MI_argmax0"])
    with pytest.raises(AssertionError):
        spectral.FM_to_p2p(g["C_closed_form"], g["Phi1"][:, :5], g["Phi2"], A1)
    p = g["ref_cf_p2p_21"]
    assert relF(spectral.p2p_to_FM(p, g["Phi1"], g["Phi2"], A2=g["area2"]), g["ref_cf_C_area"]) < 1e-12
    assert relF(spectral.p2p_to_FM(p, g["Phi1"], g["Phi2"], A2=sp.diags(g["area2"]).tocsc()), g["ref_cf_C_area"]) < 1e-12
    assert relF(spectral.p2p_to_FM(p, g["Phi1"], g["Phi2"]), g["ref_cf_C_lstsq"]) < 1e-9
    with pytest.raises(ValueError):
        spectral.p2p_to_FM(p, g["Phi1"], g["Phi2"], A2=g["area2"][:-1])


def test_zoomout_upstream_semantics(golden_zo):
    from densematcher_b200.pyFM import refine
    g = golden_zo
    C, p = refine.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=14, step=1, A2=g["area2"], return_p2p=True)
    assert C.shape == (26, 26) and relF(C, g["ref_C_zo"]) < 1e-11
    assert np.array_equal(p, g["ref_p2p_zo"])
    C, p = refine.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=9, step=(2, 3), A2=g["area2"], return_p2p=True)
    assert C.shape == (39, 30) and relF(C, g["ref_C_zo_rect"]) < 1e-11
    assert np.array_equal(p, g["ref_p2p_zo_rect"])
    C = refine.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=6, step=1, A2=g["area2"],
                              subsample=(g["sub1"], g["sub2"]))
    assert relF(C, g["ref_C_zo_sub"]) < 1e-8
    with pytest.raises(AssertionError):
        refine.zoomout_refine(g["C0"], g["Phi1"], g["Phi2"], nit=40, step=1, A2=g["area2"])


def test_batched_pairs_equal_single_pairs(fm):
    """A ragged batch of 3 pairs through one call == three single-pair oracle runs (C and all index maps)."""
    rng = np.random.default_rng(42)
    meshes = []
    for sub, scale in ((2, (1, 1.2, 0.8)), (3, (1.1, 0.9, 1.0)), (2, (0.9, 1.0, 1.3))):
        V, F = meshgen.icosphere(sub)
        meshes.append(meshgen.lbo_basis(meshgen.deform(V, scale, bump=0.1, phase=(0.2, 0.7)), F, 24))
    pairs = [(0, 1), (1, 2), (2, 0)]
    k, d = 12, 40
    Phi1 = np.concatenate([meshes[a][1] for a, _ in pairs]); Phi2 = np.concatenate([meshes[b][1] for _, b in pairs])
    ar1 = np.concatenate([meshes[a][2] for a, _ in pairs]); ar2 = np.concatenate([meshes[b][2] for _, b in pairs])
    o1 = np.concatenate([[0], np.cumsum([meshes[a][1].shape[0] for a, _ in pairs])])
    o2 = np.concatenate([[0], np.cumsum([meshes[b][1].shape[0] for _, b in pairs])])
    F1 = meshgen.random_unit_features(int(o1[-1]), d, rng); F2 = meshgen.random_unit_features(int(o2[-1]), d, rng)
    A = fm.project(dev(Phi1), dev(ar1), dev(F1), o1, k=k)
    B = fm.project(dev(Phi2), dev(ar2), dev(F2), o2, k=k)
    ev1 = np.stack([meshes[a][0][:k] for a, _ in pairs]); ev2 = np.stack([meshes[b][0][:k] for _, b in pairs])
    c00 = np.array([orc.fmap_c00(meshes[a][1], meshes[b][1], meshes[a][2], meshes[b][2]) for a, b in pairs])
    C = fm.fmap_solve(A, B, dev(ev1), dev(ev2), dev(c00), 1e4, 1e3)
    out = fm.fm_to_p2p(C, dev(Phi1[:, :k]), dev(Phi2[:, :k]), dev(ar1), o1, o2)
    Cz, pz = fm.zoomout(C, dev(Phi1), dev(Phi2), dev(ar2), nit=8, step=1, off1=o1, off2=o2, return_p2p=True)
    for i, (a, b) in enumerate(pairs):
        s1, s2 = slice(o1[i], o1[i + 1]), slice(o2[i], o2[i + 1])
        Ao, Bo = orc.project(Phi1[s1], ar1[s1], F1[s1], k), orc.project(Phi2[s2], ar2[s2], F2[s2], k)
        Co = orc.fmap_solve_closed_form(Ao, Bo, ev1[i], ev2[i], c00[i], 1e4, 1e3)
        assert relF(C[i].cpu().numpy(), Co) < 1e-4
        Cg = C[i].cpu().numpy()
        p21, p12, MI = orc.fm_to_p2p(Cg, Phi1[s1, :k], Phi2[s2, :k], ar1[s1])
        assert np.array_equal(out["p2p_21"][s2].cpu().numpy(), p21)
        assert np.array_equal(out["p2p_12"][s1].cpu().numpy(), p12)
        d21, d12 = orc.dense_argmax_override(MI)
        assert np.array_equal(out["dense_21"][s2].cpu().numpy(), d21)
        assert np.array_equal(out["dense_12"][s1].cpu().numpy(), d12)
        Czo, pzo = orc.zoomout_refine(Cg, Phi1[s1], Phi2[s2], nit=8, step=1, A2=ar2[s2], return_p2p=True)
        assert relF(Cz[i].cpu().numpy(), Czo) < 1e-10
        assert np.array_equal(pz[s2].cpu().numpy(), pzo)


def test_icp_matches_reference_golden(fm, golden_fm):
    """Spectral ICP, 10 iterations from the closed-form C: the reference's own icp_refine output
    (icp.py:43-107 run in the authoring container) is reproduced -- C within 1e-4 (in practice ~1e-12), p2p exact."""
    from densematcher_b200.pyFM import refine
    g = golden_fm
    C, p = refine.icp_refine(g["C_closed_form"], g["Phi1"], g["Phi2"], g["area1"], nit=10, return_p2p=True)
    assert relF(C, g["ref_cf_C_icp"]) < 1e-9
    assert np.array_equal(p, g["ref_cf_p2p_icp"])
    assert np.allclose(C @ C.T, np.eye(C.shape[0]), atol=1e-12)        # U I V^T of a square map is orthogonal
    # every iteration count agrees with the oracle restatement, also on a rectangular (k2 != k1) map
    for nit in (1, 3):
        Co, po = orc.icp_refine(g["C_closed_form"], g["Phi1"], g["Phi2"], nit=nit, return_p2p=True)
        Cg, pg = fm.icp(dev(g["C_closed_form"]), dev(g["Phi1"]), dev(g["Phi2"]), nit=nit, return_p2p=True)
        assert relF(Cg[0].cpu().numpy(), Co) < 1e-9 and np.array_equal(pg.cpu().numpy(), po)
    for shape in ((14, 20), (20, 13)):
        C0 = g["C_closed_form"][:shape[0], :shape[1]]
        Co, po = orc.icp_refine(C0, g["Phi1"], g["Phi2"], nit=4, return_p2p=True)
        Cg, pg = fm.icp(dev(C0), dev(g["Phi1"][:, :shape[1]]), dev(g["Phi2"][:, :shape[0]]), nit=4, return_p2p=True)
        assert relF(Cg[0].cpu().numpy(), Co) < 1e-9 and np.array_equal(pg.cpu().numpy(), po)


def test_icp_tolerance_mode_and_batch(fm, golden_fm, golden_zo):
    from densematcher_b200.pyFM import refine
    g = golden_fm
    Co = orc.icp_refine(g["C_closed_form"], g["Phi1"], g["Phi2"], nit=None, tol=1e-10)
    C = refine.icp_refine(g["C_closed_form"], g["Phi1"], g["Phi2"], g["area1"], nit=None, tol=1e-10)
    assert relF(C, Co) < 1e-8
    # two different pairs in one ragged batch == two single-pair oracle runs
    z = golden_zo
    k = 12
    Phi1 = np.concatenate([g["Phi1"][:, :k], z["Phi1"][:, :k]]); Phi2 = np.concatenate([g["Phi2"][:, :k], z["Phi2"][:, :k]])
    off = np.array([0, 642, 1284])
    C0 = np.stack([g["C_closed_form"][:k, :k], z["C0"]])
    Cg, pg = fm.icp(dev(C0), dev(Phi1), dev(Phi2), nit=5, off1=off, off2=off, return_p2p=True)
    for i, src in enumerate((g, z)):
        Co, po = orc.icp_refine(C0[i], src["Phi1"][:, :k], src["Phi2"][:, :k], nit=5, return_p2p=True)
        assert relF(Cg[i].cpu().numpy(), Co) < 1e-9
        assert np.array_equal(pg[off[i]:off[i + 1]].cpu().numpy(), po)


def test_polar_and_gram_inverse_large_k(fm):
    """k = 130 exceeds the shared-memory budget of the Jacobi kernel: the L2-scratch variant must agree too."""
    rng = np.random.default_rng(5)
    n, k = 900, 130
    Q1 = np.linalg.qr(rng.standard_normal((n, k)))[0] * (1 + 0.3 * rng.random((n, 1)))
    Q2 = np.linalg.qr(rng.standard_normal((n, k)))[0] * (1 + 0.3 * rng.random((n, 1)))
    C0 = np.linalg.qr(rng.standard_normal((k, k)))[0]
    Co, po = orc.icp_refine(C0, Q1, Q2, nit=2, return_p2p=True)
    Cg, pg = fm.icp(dev(C0), dev(Q1), dev(Q2), nit=2, return_p2p=True)
    assert relF(Cg[0].cpu().numpy(), Co) < 1e-9 and np.array_equal(pg.cpu().numpy(), po)


def test_host_entry_equals_device_pipeline():
    """match_pairs_host (pinned H2D -> staged chunks on three streams -> D2H) returns exactly what the
    device-resident pipeline returns, for a ragged batch and for chunk sizes that do / do not divide it."""
    from densematcher_b200 import pipeline
    rng = np.random.default_rng(11)
    P, d, K = 7, 48, 16
    n1 = rng.integers(150, 260, size=P); n2 = rng.integers(150, 260, size=P)
    o1 = np.concatenate([[0], np.cumsum(n1)]).astype(np.int64); o2 = np.concatenate([[0], np.cumsum(n2)]).astype(np.int64)
    b1 = [meshgen.synthetic_basis(int(n), K, rng) for n in n1]; b2 = [meshgen.synthetic_basis(int(n), K, rng) for n in n2]
    host = pipeline.PairBatchHost(
        F1=meshgen.random_unit_features(int(o1[-1]), d, rng), F2=meshgen.random_unit_features(int(o2[-1]), d, rng),
        off1=o1, off2=o2, Phi1=np.concatenate([b[1] for b in b1]), Phi2=np.concatenate([b[1] for b in b2]),
        evals1=np.stack([b[0] for b in b1]), evals2=np.stack([b[0] for b in b2]),
        area1=np.concatenate([b[2] for b in b1]), area2=np.concatenate([b[2] for b in b2])).pin()
    ref = {n: t.cpu().numpy() for n, t in pipeline.match_pairs_device(host.to_device("cuda:0"), k=12).items()}
    # the single-call path (dm_match_pairs, shared feature splits) and the stage-by-stage path agree bit for bit
    staged = pipeline.match_pairs_device(host.to_device("cuda:0"), k=12, fused=False)
    assert set(staged) == set(ref)
    for n in ref:
        assert np.array_equal(staged[n].cpu().numpy(), ref[n]), n
    for chunk in (3, 7, 64):
        for _ in range(2):  # second call reuses every staging buffer
            out = pipeline.match_pairs_host(host, "cuda:0", chunk_pairs=chunk, k=12, copy=(chunk != 7))
            assert set(out) == set(ref)
            for n in ref:
                assert out[n].dtype == ref[n].dtype and np.array_equal(out[n], ref[n]), (n, chunk)
    # and against the oracle for one pair of the batch
    s1, s2 = slice(o1[2], o1[3]), slice(o2[2], o2[3])
    assert np.array_equal(ref["nn_p2p_21"][s2], orc.nn_argmax(host.F2[s2], host.F1[s1]))
    assert np.array_equal(ref["nn_p2p_12"][s1], orc.nn_argmax(host.F2[s2], host.F1[s1], axis=0))


def test_full_size_reference_golden_fm_to_p2p_p2p_to_fm_icp(fm, golden_full):
    """BASELINE size against the REFERENCE (fm_full_ico4: icosphere(4), 2562 vertices, k = 100): all four index maps of
    FM_to_p2p + dense argmax, p2p_to_FM, and the 10-iteration icp_refine, on the reference's own inputs."""
    g = golden_full
    k = int(g["k"])
    P1, P2 = dev(g["Phi1"][:, :k]), dev(g["Phi2"][:, :k])
    C = dev(g["C_closed_form"])[None]
    out = fm.fm_to_p2p(C, P1, P2, dev(g["area1"]))
    assert np.array_equal(out["p2p_21"].cpu().numpy(), g["ref_p2p_21"])
    assert np.array_equal(out["p2p_12"].cpu().numpy(), g["ref_p2p_12"])
    assert np.array_equal(out["dense_21"].cpu().numpy(), g["ref_MI_argmax1"])
    assert np.array_equal(out["dense_12"].cpu().numpy(), g["ref_MI_argmax0"])
    Ca = fm.p2p_to_fm(dev(g["ref_p2p_21"]), P1, P2, dev(g["area2"]))[0].cpu().numpy()
    assert relF(Ca, g["ref_C_area"]) < 1e-12
    Ci, pi = fm.icp(C, P1, P2, nit=10, return_p2p=True)
    assert relF(Ci[0].cpu().numpy(), g["ref_C_icp"]) < 1e-9
    assert np.array_equal(pi.cpu().numpy(), g["ref_p2p_icp"])
    # projection + closed-form solve from the stored descriptors land on the stored C (tcgen05 projection: fp32-grade)
    A = fm.project(P1, dev(g["area1"]), dev(g["c1"]), k=k)
    B = fm.project(P2, dev(g["area2"]), dev(g["c2"]), k=k)
    c00 = orc.fmap_c00(g["Phi1"], g["Phi2"], g["area1"], g["area2"])
    Cs = fm.fmap_solve(A, B, dev(g["evals1"][:k])[None], dev(g["evals2"][:k])[None], dev(np.array([c00])),
                       float(g["w_descr"]), float(g["w_lap"]))[0].cpu().numpy()
    assert relF(Cs, g["C_closed_form"]) < 1e-4          # north-star bar; ~1e-6 in practice


def test_full_size_reference_golden_zoomout_ladder(fm, golden_full):
    """ZoomOut 30 -> 200 (170 rungs) at N = 2562 against the ladder composed from the REFERENCE's knn_query / p2p_to_FM:
    final C and final p2p; the default (float64 C) mode must reproduce them exactly."""
    g = golden_full
    C0 = dev(g["C_closed_form"][:30, :30].copy())[None]
    Cz, pz = fm.zoomout(C0, dev(g["Phi1"]), dev(g["Phi2"]), dev(g["area2"]), nit=170, step=1, return_p2p=True)
    assert relF(Cz[0].cpu().numpy(), g["ref_C_zo"]) < 1e-10
    assert np.array_equal(pz.cpu().numpy(), g["ref_p2p_zo"])


def test_zoomout_ladder_full_size_30_to_200():
    """BASELINE config 4 shape for one pair: N = 2000, ladder k = 30 -> 200 step 1 (170 iterations).  Every
    intermediate p2p must match for the final C to match: C within 1e-4 (in practice ~1e-10) and
    the final map bit-exact against the float64 oracle."""
    rng = np.random.default_rng(4000)
    n, K = 2000, 200
    e1, P1, a1 = meshgen.synthetic_basis(n, K, rng)
    e2, P2, a2 = meshgen.synthetic_basis(n, K, rng)
    # a smooth-ish ground-truth relation between the bases so that the ladder is not pure noise
    C0 = np.linalg.qr(rng.standard_normal((30, 30)))[0]
    Cz, pz = fm_mod().zoomout(dev(C0), dev(P1), dev(P2), dev(a2), nit=170, step=1, return_p2p=True)
    Co, po = orc.zoomout_refine(C0, P1, P2, nit=170, step=1, A2=a2, return_p2p=True)
    assert Cz.shape == (1, 200, 200)
    assert relF(Cz[0].cpu().numpy(), Co) < 1e-4
    assert np.array_equal(pz.cpu().numpy(), po)


def fm_mod():
    from densematcher_b200 import fm as _fm
    return _fm


def test_zoomout_fast_mode_stays_close(golden_zo):
    """DM_FAST_FM (tensor-core accumulation of C, fp32-grade) is an opt-in: C within 1e-3 of the float64 ladder and
    the final map equal on all but a handful of near-tie vertices."""
    from densematcher_b200 import _lib
    g = golden_zo
    C, p = fm_mod().zoomout(dev(g["C0"]), dev(g["Phi1"]), dev(g["Phi2"]), dev(g["area2"]), nit=14, step=1,
                            return_p2p=True, flags=_lib.DM_FAST_FM)
    assert relF(C[0].cpu().numpy(), g["ref_C_zo"]) < 1e-3
    assert np.mean(p.cpu().numpy() != g["ref_p2p_zo"]) < 0.01


def test_diffusionnet_to_basis_matches_torch():
    """to_basis (diffusion_net/geometry.py:572-583) on the tcgen05 projection engine vs the plain float64 product."""
    import torch
    from densematcher_b200.spectral_ops import to_basis, from_basis
    g = torch.Generator(device="cuda").manual_seed(7)
    B, V, K, D = 3, 700, 64, 96
    vals = torch.randn(B, V, D, device="cuda", generator=g)
    basis = torch.linalg.qr(torch.randn(B, V, K, device="cuda", generator=g, dtype=torch.float64))[0]
    mass = torch.rand(B, V, device="cuda", generator=g, dtype=torch.float64) + 0.5
    ref = basis.transpose(1, 2) @ (vals.double() * mass[..., None])
    out = to_basis(vals, basis, mass)
    assert out.shape == (B, K, D) and float((out.double() - ref).norm() / ref.norm()) < 5e-6
    back = from_basis(out[0].double(), basis[0])
    assert back.shape == (V, D)


@pytest.mark.parametrize("k1,k2", [(2, 2), (3, 5), (33, 20), (64, 64), (129, 100), (160, 170), (236, 30)])
def test_fmap_solve_sizes(k1, k2):
    """Closed-form solve across system sizes: tiny, non-square, one row per thread (k1 <= 128) and two (k1 > 128),
    up to the shared-memory limit (k1 = 236)."""
    rng = np.random.default_rng(k1 * 1000 + k2)
    d, P = 48, 2
    A = rng.standard_normal((P, k1, d)); B = rng.standard_normal((P, k2, d))
    ev1 = np.sort(rng.random((P, k1)) * 50, axis=1); ev2 = np.sort(rng.random((P, k2)) * 50, axis=1)
    ev1[:, 0] = ev2[:, 0] = 0.0
    c00 = rng.standard_normal(P)
    C, st = fm_mod().fmap_solve(dev(A), dev(B), dev(ev1), dev(ev2), dev(c00), 3.0, 0.7, return_status=True)
    C = C.cpu().numpy()
    assert st[0] == 0
    # d = 48 < k1 - 1 makes the Gram part rank deficient for the larger sizes: those systems are regularised only by the
    # Laplacian diagonal, the float32 preconditioner does not contract on all of them and the float64 kernel takes over
    # (st[1] counts them); either way the result is the float64 solution.  Bound: 1e-9 relative Frobenius.
    for p in range(P):
        Co = orc.fmap_solve_closed_form(A[p], B[p], ev1[p], ev2[p], c00[p], 3.0, 0.7)
        assert relF(C[p], Co) < 1e-9, (k1, k2, p, st)
    if k1 == 236:
        with pytest.raises(Exception):
            fm_mod().fmap_solve(dev(rng.standard_normal((1, 260, 8))), dev(B[:1, :, :8]), dev(np.zeros((1, 260))),
                                dev(ev2[:1]), dev(c00[:1]), 1.0, 1.0)


def test_fmap_solve_float32_refinement_reaches_float64(monkeypatch):
    """The default solve (float32 Cholesky as a preconditioner + float64 refinement against the float64 Gram matrix) at
    the bench shape: every system converges without the float64 fallback, in one or two refinement steps, to the same
    C as the float64 kernel (DM_SOLVE=f64) and the 64-thread float32 variant (DM_SOLVE=f32t64)."""
    rng = np.random.default_rng(5)
    P, k, d = 3, 100, 384
    A = rng.standard_normal((P, k, d)) * 0.05; B = rng.standard_normal((P, k, d)) * 0.05
    ev = np.sort(rng.random((P, k)) * 80, axis=1); ev[:, 0] = 0.0
    ev2 = np.sort(rng.random((P, k)) * 80, axis=1); ev2[:, 0] = 0.0
    c00 = np.array([1.0, -1.0, 0.9])
    args = (dev(A), dev(B), dev(ev), dev(ev2), dev(c00), 1e4, 1e3)
    C, st = fm_mod().fmap_solve(*args, return_status=True)
    C = C.cpu().numpy()
    assert st[0] == 0 and st[1] == 0, st                 # nothing singular, nothing sent to the fallback
    assert P * k <= st[3] <= 2 * P * k, st               # refinement steps per system
    for p in range(P):
        Co = orc.fmap_solve_closed_form(A[p], B[p], ev[p], ev2[p], c00[p], 1e4, 1e3)
        assert relF(C[p], Co) < 1e-10, p
    for mode in ("f64", "f32t64"):
        monkeypatch.setenv("DM_SOLVE", mode)
        C2 = fm_mod().fmap_solve(*args).cpu().numpy()
        assert relF(C2, C) < 1e-10, mode
    monkeypatch.delenv("DM_SOLVE")


def test_fmap_solve_reports_singular_systems():
    """w_lap = 0 with fewer descriptors than unknowns: the row systems are singular.  The reference's L-BFGS returns some
    finite map there; the closed form cannot, and must say so instead of returning NaN silently."""
    from densematcher_b200._lib import DMError
    rng = np.random.default_rng(6)
    P, k, d = 2, 24, 8
    A = rng.standard_normal((P, k, d)); B = rng.standard_normal((P, k, d))
    ev = np.sort(rng.random((P, k)), axis=1)
    args = (dev(A), dev(B), dev(ev), dev(ev), dev(np.ones(P)), 1.0, 0.0)
    with pytest.raises(DMError):
        fm_mod().fmap_solve(*args)
    C, st = fm_mod().fmap_solve(*args, check=False, return_status=True)
    assert st[0] != 0 and st[1] > 0


def test_random_shape_sweep_projection_p2p_to_fm_fm_to_p2p():
    """Randomised shapes through the tensor-core projection, the float64 GEMM paths and the fused FM->p2p pass."""
    rng = np.random.default_rng(99)
    f = fm_mod()
    for trial in range(12):
        P = int(rng.integers(1, 4))
        n1, n2 = rng.integers(3, 400, size=P), rng.integers(3, 400, size=P)
        k1, k2 = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        k1, k2 = min(k1, int(n1.min())), min(k2, int(n2.min()))
        d = int(rng.integers(1, 90))
        o1 = np.concatenate([[0], np.cumsum(n1)]); o2 = np.concatenate([[0], np.cumsum(n2)])
        Phi1 = rng.standard_normal((o1[-1], k1)); Phi2 = rng.standard_normal((o2[-1], k2))
        a1 = rng.random(o1[-1]) + 0.1; a2 = rng.random(o2[-1]) + 0.1
        F1 = rng.standard_normal((o1[-1], d)).astype(np.float32)
        A = f.project(dev(Phi1), dev(a1), dev(F1), o1).cpu().numpy()
        C = rng.standard_normal((P, k2, k1))
        out = f.fm_to_p2p(dev(C), dev(Phi1), dev(Phi2), dev(a1), o1, o2)
        p21 = out["p2p_21"].cpu().numpy()
        Cn = f.p2p_to_fm(dev(p21), dev(Phi1), dev(Phi2), dev(a2), o1, o2).cpu().numpy()
        for p in range(P):
            s1, s2 = slice(o1[p], o1[p + 1]), slice(o2[p], o2[p + 1])
            ref = orc.project(Phi1[s1], a1[s1], F1[s1])
            assert np.abs(A[p] - ref).max() < 5e-6 * np.abs(ref).max() + 1e-9, (trial, "project")
            r21, r12, MI = orc.fm_to_p2p(C[p], Phi1[s1], Phi2[s2], a1[s1])
            assert np.array_equal(p21[s2], r21) and np.array_equal(out["p2p_12"][s1].cpu().numpy(), r12), (trial, "p2p")
            assert np.array_equal(out["dense_21"][s2].cpu().numpy(), MI.argmax(1)), (trial, "dense_21")
            assert np.array_equal(out["dense_12"][s1].cpu().numpy(), MI.argmax(0)), (trial, "dense_12")
            assert relF(Cn[p], orc.p2p_to_fm(r21, Phi1[s1], Phi2[s2], a2[s2])) < 1e-11, (trial, "p2p_to_fm")


def test_polar_factor_newton_schulz_and_jacobi_fallback():
    """U I V^T against scipy's SVD: well-conditioned matrices (Newton-Schulz path), an ill-conditioned and a
    rank-deficient one in the same batch (Jacobi fallback), tall and wide shapes, and the forced Jacobi path."""
    import scipy.linalg
    from densematcher_b200 import _lib
    rng = np.random.default_rng(17)

    def ref(X):
        U, _, Vt = scipy.linalg.svd(X)
        return U @ np.eye(*X.shape) @ Vt

    for rows, cols in ((40, 40), (100, 100), (30, 22), (22, 30), (130, 130)):
        k = min(rows, cols)
        Xs = []
        for cond in (1.5, 5.0, 1e6):
            U = np.linalg.qr(rng.standard_normal((rows, rows)))[0][:, :k]
            V = np.linalg.qr(rng.standard_normal((cols, cols)))[0][:, :k]
            Xs.append(U @ np.diag(np.geomspace(1.0, 1.0 / cond, k)) @ V.T * rng.uniform(0.2, 5))
        X = np.stack(Xs)
        for flags in (0, _lib.DM_POLAR_JACOBI):
            C = fm_mod().polar_factor(dev(X), flags=flags).cpu().numpy()
            for b in range(3):
                tol = 1e-11 if b < 2 else 1e-7      # cond 1e6: the polar factor itself is that sensitive
                assert np.abs(C[b] - ref(X[b])).max() < tol, (rows, cols, b, flags)
                G = C[b].T @ C[b] if rows >= cols else C[b] @ C[b].T
                assert np.abs(G - np.eye(k)).max() < 1e-11, (rows, cols, b, flags)


def test_float64_gemm_paths_with_unaligned_operands(fm):
    """The float64 GEMMs move operands 16 bytes at a time when rows start on 16-byte boundaries and fall back to
    8-byte accesses otherwise: odd leading dimensions, odd sizes and column-offset views must give the same numbers."""
    rng = np.random.default_rng(31)
    n1, n2 = 700, 650
    for (k1, k2, ld, c0) in ((21, 19, 45, 1), (20, 20, 44, 0), (33, 30, 47, 2), (104, 100, 104, 0)):
        big1, big2 = rng.standard_normal((n1, ld + 3)), rng.standard_normal((n2, ld + 3))
        C = rng.standard_normal((k2, k1))
        a1 = rng.uniform(0.5, 1.5, n1)
        P1, P2 = big1[:, c0:c0 + k1], big2[:, c0:c0 + k2]
        MI = fm.mapped_indicator(dev(C), dev(big1)[:, c0:c0 + k1], dev(big2)[:, c0:c0 + k2], dev(a1)).cpu().numpy()
        ref = (P2 @ C @ P1.T) * a1[None, :]
        assert np.abs(MI - ref).max() < 1e-10 * np.abs(ref).max(), (k1, k2, ld, c0)
        # p2p -> FM (transposed, gathered, scaled operands) on the same views
        p = rng.integers(0, n1, n2)
        a2 = rng.uniform(0.5, 1.5, n2)
        Cg = fm.p2p_to_fm(dev(p), dev(big1)[:, c0:c0 + k1], dev(big2)[:, c0:c0 + k2], dev(a2))[0].cpu().numpy()
        assert relF(Cg, P2.T @ (a2[:, None] * P1[p])) < 1e-12, (k1, k2, ld, c0)


@pytest.mark.gpu
@pytest.mark.parametrize("step", [1, (2, 1), (1, 3), (4, 4)])
def test_zoomout_incremental_rungs_ragged_rectangular(fm, step):
    """The ladder keeps M = Phi2^T A2 Phi1[p] resident and corrects it with the changed vertices (zoomout_delta.cu): a
    ragged batch (three pairs of different sizes), square and rectangular steps, int32 and int64 maps, long enough to
    cross a re-anchoring rung -- every pair against the float64 oracle ladder (final p2p identical, C to 1e-10), i.e.
    the same result as the full product per rung."""
    rng = np.random.default_rng(77)
    meshes = []
    for sub, scale in ((2, (1, 1.2, 0.8)), (3, (1.1, 0.9, 1.0)), (2, (0.9, 1.0, 1.3))):
        V, F = meshgen.icosphere(sub)
        meshes.append(meshgen.lbo_basis(meshgen.deform(V, scale, bump=0.1, phase=(0.2, 0.7)), F, 150))
    pairs = [(0, 1), (1, 2), (2, 0)]
    s1, s2 = (step, step) if isinstance(step, int) else step
    k0 = 6
    nit = min((150 - k0) // max(s1, s2), 70)                      # 70 rungs at step 1: crosses the anchor at rung 64
    Phi1 = np.concatenate([meshes[a][1] for a, _ in pairs]); Phi2 = np.concatenate([meshes[b][1] for _, b in pairs])
    ar2 = np.concatenate([meshes[b][2] for _, b in pairs])
    o1 = np.concatenate([[0], np.cumsum([meshes[a][1].shape[0] for a, _ in pairs])])
    o2 = np.concatenate([[0], np.cumsum([meshes[b][1].shape[0] for _, b in pairs])])
    C0 = np.stack([np.linalg.qr(rng.standard_normal((k0, k0)))[0] for _ in pairs])
    for dt in (torch.int32, torch.int64):
        Cz, pz = fm.zoomout(dev(C0), dev(Phi1), dev(Phi2), dev(ar2), nit=nit, step=step, off1=o1, off2=o2, return_p2p=True,
                            out_dtype=dt)
        assert pz.dtype == dt
        for i in range(len(pairs)):
            a, b = slice(o1[i], o1[i + 1]), slice(o2[i], o2[i + 1])
            Co, po = orc.zoomout_refine(C0[i], Phi1[a], Phi2[b], nit=nit, step=step, A2=ar2[b], return_p2p=True)
            assert np.array_equal(pz[b].cpu().numpy(), po), (i, step)
            assert relF(Cz[i].cpu().numpy(), Co) < 1e-10
