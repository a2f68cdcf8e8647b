"""The drop-in driver: ``compute_surface_map`` / ``FunctionalMapping`` against what the reference's own
``compute_surface_map`` produced in the authoring container (tests/golden/fm_pair_ico3.npz)."""
import numpy as np
import pytest

from oracle import dm_oracle as orc

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


class _Mesh:  # pytorch3d-like duck: only verts_list()/faces_list() are read (functional_map.py:17-18)
    def __init__(self, V, F):
        import torch
        self._v, self._f = torch.tensor(V, dtype=torch.float32), torch.tensor(F, dtype=torch.int64)

    def verts_list(self):
        return [self._v]

    def faces_list(self):
        return [self._f]


def _meshes(g):
    from densematcher_b200.pyFM import TriMesh
    return (TriMesh.from_basis(g["evals1"], g["Phi1"], g["area1"]), TriMesh.from_basis(g["evals2"], g["Phi2"], g["area2"]))


def test_compute_surface_map_matches_reference_outputs(golden_fm):
    from densematcher_b200 import _lib
    from densematcher_b200.functional_map import compute_surface_map
    from densematcher_b200.pyFM import FunctionalMapping
    g = golden_fm
    k = int(g["k"])
    fit = dict(w_descr=float(g["w_descr"]), w_lap=float(g["w_lap"]), w_dcomm=0)
    m1, m2 = _meshes(g)
    # (1) float64 projection: every array the reference computed FROM THE CLOSED-FORM C is reproduced exactly
    FunctionalMapping.projection_flags = _lib.DM_F64_GEMM
    try:
        out = compute_surface_map(m1, m2, g["c1"], g["c2"], n_ev=k, fit_params=fit)
    finally:
        FunctionalMapping.projection_flags = 0
    assert len(out) == 14
    model = out[7]
    assert relF(model._FM_base, g["C_closed_form"]) < 1e-9
    assert np.array_equal(out[10], g["ref_cf_p2p_21"]) and np.array_equal(out[11], g["ref_cf_p2p_12"])
    assert np.array_equal(out[0], g["ref_cf_MI_argmax1"]) and np.array_equal(out[1], g["ref_cf_MI_argmax0"])
    assert relF(model._FM_icp, g["ref_cf_C_icp"]) < 1e-9 and model.FM_type == "icp"
    assert np.array_equal(out[12], g["ref_cf_p2p_icp"])
    assert out[0].dtype == np.int64 and out[3] is None and out[2] is None
    # Hungarian slot (host scipy on the materialised mapped_indicator): a full assignment of the 642 vertices
    rows, cols = out[6]
    assert np.array_equal(rows, np.arange(642)) and len(set(cols.tolist())) == 642
    MI = model.mapped_indicator
    r21, r12, MIo = orc.fm_to_p2p(model.FM, g["Phi1"], g["Phi2"], g["area1"])
    assert np.allclose(MI, MIo, rtol=1e-10, atol=1e-13) and np.array_equal(out[4], MIo.argmax(1))
    # (2) default tensor-core projection: C within the 1e-4 bar; the index maps agree with the reference's own
    #     end-to-end run (L-BFGS C) as well as the reference agrees with itself (a few vertices)
    out2 = compute_surface_map(m1, m2, g["c1"], g["c2"], n_ev=k, fit_params=fit, hungarian=False)
    assert relF(out2[7]._FM_base, g["C_closed_form"]) < 1e-4 and out2[6] is None
    for slot, name in ((0, "ref_p2p_21"), (1, "ref_p2p_12"), (10, "ref_p2p_21_adjoint"), (11, "ref_p2p_12_adjoint")):
        assert np.mean(out2[slot] != g[name]) < 0.02, name
    assert np.mean(out2[0] != out[0]) < 0.01


def test_functional_mapping_surface_and_errors(golden_fm, golden_zo):
    from densematcher_b200.pyFM import FunctionalMapping
    g = golden_fm
    k = int(g["k"])
    m1, m2 = _meshes(g)
    model = FunctionalMapping(m1, m2, partial=False)
    with pytest.raises(ValueError):
        _ = model.k1
    model.preprocess(n_ev=(k, k), descr1=g["c1"], descr2=g["c2"])
    assert model.preprocessed and not model.fitted and (model.k1, model.k2) == (k, k)
    with pytest.raises(NotImplementedError):
        model.fit(w_descr=1e4, w_lap=1e3, w_dcomm=0, w_orient=1.0)     # unimplemented energy term: never ignored
    with pytest.raises(ValueError):
        model.get_p2p()
    model.fit(w_descr=float(g["w_descr"]), w_lap=float(g["w_lap"]), w_dcomm=0)
    assert model.fitted and model.FM.shape == (k, k) and np.all(model.eta == 1)
    with pytest.raises(ValueError):
        model.FM_type = "bogus"
    # transfer of a band-limited function = decode(C @ project(f))  (functional.py:806-831)
    f = g["Phi1"][:, :k] @ np.random.default_rng(0).standard_normal((k, 3))
    enc = model.project(f)
    assert np.allclose(enc, orc.project(g["Phi1"], g["area1"], f, k), rtol=1e-10, atol=1e-12)
    assert np.allclose(model.transfer(f), g["Phi2"][:, :k] @ (model.FM @ enc), rtol=1e-9, atol=1e-12)
    # zoomout_refine through the class == the free function (upstream semantics)
    z = golden_zo
    from densematcher_b200.pyFM import TriMesh
    zm = FunctionalMapping(TriMesh.from_basis(z["evals1"], z["Phi1"], z["area1"]),
                           TriMesh.from_basis(z["evals2"], z["Phi2"], z["area2"]))
    zm.FM = z["C0"]
    zm.zoomout_refine(nit=14, step=1)
    assert zm.FM_type == "zoomout" and relF(zm.FM, z["ref_C_zo"]) < 1e-11


def test_trimesh_host_spectrum_fallback():
    """Bare geometry in (pytorch3d-like duck): the host LBO fallback produces an A-orthonormal basis and the driver
    runs end to end."""
    from oracle import meshgen
    from densematcher_b200.functional_map import compute_surface_map
    V, F = meshgen.icosphere(2)
    V2 = meshgen.deform(V, (1.1, 0.9, 1.0), bump=0.1)
    rng = np.random.default_rng(3)
    c1 = meshgen.random_unit_features(V.shape[0], 24, rng)
    out = compute_surface_map(_Mesh(V, F), _Mesh(V2, F), c1, c1.copy(), n_ev=10,
                              fit_params=dict(w_descr=1e4, w_lap=1e3, w_dcomm=0), hungarian=False)
    m1 = out[8]
    G = m1.eigenvectors.T @ (m1.vertex_areas[:, None] * m1.eigenvectors)
    assert np.allclose(G, np.eye(10), atol=1e-8) and abs(m1.eigenvalues[0]) < 1e-6
    # the maps are those of the oracle for the same (host-computed) bases and the model's C
    m2, model = out[9], out[7]
    r21, r12, MI = orc.fm_to_p2p(model._FM_base, m1.eigenvectors, m2.eigenvectors, m1.vertex_areas)
    assert np.array_equal(out[10], r21) and np.array_equal(out[11], r12)
    assert np.array_equal(out[0], MI.argmax(1)) and np.array_equal(out[1], MI.argmax(0))


def test_mesh_bank_intra_category_pairs_cfg5_shape():
    """BASELINE config 5 stand-in at reduced scale: a bank of meshes in a few categories, every ordered
    intra-category pair matched from the device-resident bank; equals the packed-batch pipeline and the oracle."""
    import torch
    from oracle import meshgen
    from densematcher_b200 import pipeline
    rng = np.random.default_rng(5000)
    cats = np.array([0, 0, 0, 1, 1, 2, 2, 2, 2])
    sizes = rng.integers(180, 260, size=len(cats))
    K, d, k = 16, 48, 12
    bases = [meshgen.synthetic_basis(int(n), K, rng) for n in sizes]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    F = meshgen.random_unit_features(int(off[-1]), d, rng)
    bank = pipeline.MeshBankDevice(F, off, Phi=np.concatenate([b[1] for b in bases]), evals=np.stack([b[0] for b in bases]),
                                   area=np.concatenate([b[2] for b in bases]))
    src, dst = pipeline.intra_category_pairs(cats)
    assert len(src) == 3 * 2 + 2 * 1 + 4 * 3 and np.all(cats[src] == cats[dst]) and np.all(src != dst)
    chunks, (lo, hi) = pipeline.match_bank_pairs(bank, src, dst, chunk_pairs=7, k=k)
    assert (lo, hi) == (0, len(src)) and len(chunks) == 3
    # the same pairs as one packed host batch
    sl = lambda a, i: a[off[i]:off[i + 1]]
    host = pipeline.PairBatchHost(
        F1=np.concatenate([sl(F, i) for i in src]), F2=np.concatenate([sl(F, j) for j in dst]),
        off1=np.concatenate([[0], np.cumsum(sizes[src])]), off2=np.concatenate([[0], np.cumsum(sizes[dst])]),
        Phi1=np.concatenate([bases[i][1] for i in src]), Phi2=np.concatenate([bases[j][1] for j in dst]),
        evals1=np.stack([bases[i][0] for i in src]), evals2=np.stack([bases[j][0] for j in dst]),
        area1=np.concatenate([bases[i][2] for i in src]), area2=np.concatenate([bases[j][2] for j in dst]))
    ref = {n: t.cpu().numpy() for n, t in pipeline.match_pairs_device(host.to_device("cuda:0"), k=k).items()}
    for n in ref:
        got = np.concatenate([c[n] for c in chunks])
        assert np.array_equal(got, ref[n]), n
    # oracle on the first pair
    i, j = int(src[0]), int(dst[0])
    assert np.array_equal(chunks[0]["nn_p2p_21"][:sizes[j]], orc.nn_argmax(sl(F, j), sl(F, i)))
    # two ranks cover the list without overlap
    blocks = [pipeline.match_bank_pairs(bank, src, dst, chunk_pairs=64, rank=r, world=2, feature_nn=True,
                                        functional_map=False)[1] for r in range(2)]
    assert blocks[0][0] == 0 and blocks[0][1] == blocks[1][0] and blocks[1][1] == len(src)


def test_dense_energy_kernel_matches_oracle(golden_fm):
    """dm_dense_energy (fused tile -> loss -> contraction, M never stored) against the float64 oracle whose values are
    pinned to the reference's torch terms: every term alone, all together, a ragged two-pair batch, k1 != k2."""
    import torch
    from conftest import load_golden
    from densematcher_b200 import fm
    e = load_golden("energy_ico3.npz")
    g = golden_fm
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    k = int(e["k"])
    P1, P2, a1 = g["Phi1"][:, :k], g["Phi2"][:, :k], g["area1"]
    for name in orc.DENSE_TERMS + ("all",):
        w = {t: 0.3 + 0.1 * i for i, t in enumerate(orc.DENSE_TERMS)} if name == "all" else {name: 1.0}
        en, gr = fm.dense_energy(dev(e["C"]), dev(P1), dev(P2), dev(a1), w)
        Eo, Go, parts = orc.dense_map_energy(e["C"], P1, P2, a1, w)
        for i, t in enumerate(orc.DENSE_TERMS):
            if t in w:
                assert np.isclose(float(en[0, i]), parts[t], rtol=1e-11), (name, t)
                if name != "all":
                    assert np.isclose(float(en[0, i]), float(e["ref_E_" + t]), rtol=1e-11), t   # the reference's value
        assert np.abs(gr[0].cpu().numpy() - Go).max() < 1e-10 * np.abs(Go).max(), name
    # ragged batch of two pairs, rectangular map
    rng = np.random.default_rng(1)
    k1, k2 = 9, 13
    n1a, n2a = 400, 350
    Phi1 = np.concatenate([g["Phi1"][:n1a, :k1], g["Phi1"][:, :k1]]); Phi2 = np.concatenate([g["Phi2"][:n2a, :k2], g["Phi2"][:, :k2]])
    ar1 = np.concatenate([a1[:n1a], a1])
    o1, o2 = np.array([0, n1a, n1a + 642]), np.array([0, n2a, n2a + 642])
    C = 0.3 * rng.standard_normal((2, k2, k1))
    w = dict(ent=0.1, sumto1=10.0, stochastic=0.05, p2p=0.2, range01=1.0)
    en, gr = fm.dense_energy(dev(C), dev(Phi1), dev(Phi2), dev(ar1), w, o1, o2)
    for p in range(2):
        s1, s2 = slice(o1[p], o1[p + 1]), slice(o2[p], o2[p + 1])
        Eo, Go, parts = orc.dense_map_energy(C[p], Phi1[s1], Phi2[s2], ar1[s1], w)
        assert np.allclose(en[p].cpu().numpy(), [parts[t] for t in orc.DENSE_TERMS], rtol=1e-10)
        assert np.abs(gr[p].cpu().numpy() - Go).max() < 1e-10 * np.abs(Go).max()
    # DM_FAST_LOSS: float32 logarithm / division in the entropy term (the reference's own precision): 1e-6 of float64
    from densematcher_b200 import _lib
    en_f, gr_f = fm.dense_energy(dev(C), dev(Phi1), dev(Phi2), dev(ar1), w, o1, o2, flags=_lib.DM_FAST_LOSS)
    assert np.allclose(en_f.cpu().numpy(), en.cpu().numpy(), rtol=1e-6)
    assert float((gr_f - gr).abs().max()) < 1e-6 * float(gr.abs().max())
    # the small batched products of the fit on the library's own GEMM
    A3, B3 = rng.standard_normal((3, 13, 40)), rng.standard_normal((3, 9, 40))
    assert np.abs(fm.bmm_nt(dev(A3), dev(B3)).cpu().numpy() - A3 @ B3.transpose(0, 2, 1)).max() < 1e-12


def test_fit_with_notebook_default_weights(golden_fm):
    """example.ipynb cell 11 verbatim: w_ent = 0.1, w_sumto1 = 10 next to w_descr / w_lap, n_ev = 15, L-BFGS-B.  The
    fit lands on the reference's own result (its float32 energy leaves ~1e-4 of noise) and yields the same p2p."""
    from conftest import load_golden
    from densematcher_b200 import _lib
    from densematcher_b200.pyFM import FunctionalMapping
    e = load_golden("energy_ico3.npz")
    g = golden_fm
    k = int(e["k"])
    m1, m2 = _meshes(g)
    fit_params = {'w_descr': 1e4, 'w_lap': 1e3, 'w_dcomm': 0e0, 'w_orient': 0, 'w_area': 0, 'w_conformal': 0e1, 'w_p2p': 0,
                  'w_stochastic': 0, 'w_ent': 1e-1, 'w_range01': 0, 'w_sumto1': 1e1, 'optinit': 'zeros', 'maxiter': 5000}
    model = FunctionalMapping(m1, m2, partial=False, optimizer="L-BFGS-B")
    model.projection_flags = _lib.DM_F64_GEMM
    model.preprocess(n_ev=(k, k), descr1=g["c1"], descr2=g["c2"])
    model.fit(**fit_params)
    Cr = e["ref_C_notebook"]
    assert relF(model.FM, Cr) < 1e-3 and model.fit_result.nit > 5
    p21, p12 = model.get_p2p()
    r21, r12, _ = orc.fm_to_p2p(Cr, g["Phi1"][:, :k], g["Phi2"][:, :k], g["area1"])
    assert np.mean(p21 != r21) < 0.01 and np.mean(p12 != r12) < 0.01
    with pytest.raises(NotImplementedError):
        model.fit(w_descr=1e4, w_lap=1e3, w_dcomm=1.0)


def test_batched_device_fit_matches_reference_and_host_loop(golden_fm):
    """fm.fit_dense: the notebook's energy (w_ent = 0.1, w_sumto1 = 10 next to w_descr / w_lap) minimised for a BATCH of
    pairs by the on-device L-BFGS.  Every pair lands on the reference's own fit (< 1e-3: its float32 energy leaves
    ~1e-4 of noise) and on the host scipy loop around the same kernel (< 1e-5), with the same p2p as the reference C."""
    import torch
    from conftest import load_golden
    from densematcher_b200 import fm, _lib
    from densematcher_b200.pyFM import FunctionalMapping
    e = load_golden("energy_ico3.npz")
    g = golden_fm
    k = int(e["k"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    P1, P2, a1, a2 = g["Phi1"][:, :k], g["Phi2"][:, :k], g["area1"], g["area2"]
    A = fm.project(dev(P1), dev(a1), dev(g["c1"]), flags=_lib.DM_F64_GEMM)
    B = fm.project(dev(P2), dev(a2), dev(g["c2"]), flags=_lib.DM_F64_GEMM)
    c00 = orc.fmap_c00(g["Phi1"], g["Phi2"], a1, a2)
    n1, n2 = len(a1), len(a2)
    reps = 3   # the same pair three times, as a ragged batch
    cat = lambda x: dev(np.concatenate([x] * reps))
    off1, off2 = np.arange(reps + 1) * n1, np.arange(reps + 1) * n2
    C, info = fm.fit_dense(A.repeat(reps, 1, 1), B.repeat(reps, 1, 1), dev(np.tile(g["evals1"][:k], (reps, 1))),
                           dev(np.tile(g["evals2"][:k], (reps, 1))), dev(np.full(reps, c00)), cat(P1), cat(P2), cat(a1),
                           {"ent": 1e-1, "sumto1": 1e1}, 1e4, 1e3, off1=off1, off2=off2, return_info=True)
    C = C.cpu().numpy()
    assert 5 < info[0] < 500
    for p in range(reps):
        assert relF(C[p], e["ref_C_notebook"]) < 1e-3
        assert relF(C[p], C[0]) < 1e-9                          # batch entries do not interact
    m1, m2 = _meshes(g)
    model = FunctionalMapping(m1, m2, partial=False, optimizer="scipy")
    model.projection_flags = _lib.DM_F64_GEMM
    model.preprocess(n_ev=(k, k), descr1=g["c1"], descr2=g["c2"])
    model.fit(w_descr=1e4, w_lap=1e3, w_dcomm=0, w_ent=1e-1, w_sumto1=1e1, maxiter=5000)
    assert relF(C[0], model.FM) < 2e-4                          # two optimisers stopped by the same (scipy) tolerances
    r21, r12, _ = orc.fm_to_p2p(e["ref_C_notebook"], P1, P2, a1)
    o21, o12, _ = orc.fm_to_p2p(C[0], P1, P2, a1)
    assert np.mean(o21 != r21) < 0.01 and np.mean(o12 != r12) < 0.01


def test_cfg2_full_size_pipeline_against_oracle():
    """BASELINE config 2 at full size (N = M = 2000, d = 384, k = 100, notebook weights) for a small batch through the
    single-call pipeline: feature NN bit-exact, C within 1e-4 of the float64 closed form, and -- given that C -- all
    four FM->p2p index maps bit-exact against the float64 oracle."""
    import torch
    import bench
    from densematcher_b200 import pipeline
    P = 3
    host = bench.make_host_batch(P, seed=2222, pool=4)
    out = pipeline.match_pairs_device(host.to_device("cuda:0"), k=100, w_descr=1e4, w_lap=1e3, out_dtype=torch.int64)
    out = {n: t.cpu().numpy() for n, t in out.items()}
    for p in range(P):
        s1, s2 = slice(host.off1[p], host.off1[p + 1]), slice(host.off2[p], host.off2[p + 1])
        F1, F2, P1, P2, a1, a2 = host.F1[s1], host.F2[s2], host.Phi1[s1], host.Phi2[s2], host.area1[s1], host.area2[s2]
        assert np.array_equal(out["nn_p2p_21"][s2], orc.nn_argmax(F2, F1))
        assert np.array_equal(out["nn_p2p_12"][s1], orc.nn_argmax(F2, F1, axis=0))
        Co = orc.fmap_solve_closed_form(orc.project(P1, a1, F1), orc.project(P2, a2, F2), host.evals1[p], host.evals2[p],
                                        orc.fmap_c00(P1, P2, a1, a2), 1e4, 1e3)
        assert relF(out["C"][p], Co) < 1e-4
        r21, r12, MI = orc.fm_to_p2p(out["C"][p], P1, P2, a1)
        assert np.array_equal(out["p2p_21_adjoint"][s2], r21) and np.array_equal(out["p2p_12_adjoint"][s1], r12)
        assert np.array_equal(out["p2p_21"][s2], MI.argmax(1)) and np.array_equal(out["p2p_12"][s1], MI.argmax(0))
