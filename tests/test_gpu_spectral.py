"""GPU parity tests of the eigenbasis provider and the DiffusionNet spectral transforms.

* ``lbo_eigs`` (csrc/spectral.cu) against the REFERENCE's ``laplacian_spectrum`` outputs stored in the goldens
  (tests/golden/fm_pair_ico3.npz, fm_full_ico4.npz were minted by oracle/make_goldens.py from
  densematcher/pyFM/mesh/laplacian.py:143-182 on the same deterministic meshes), plus the defining properties
  (residual, A-orthonormality) and scipy's shift-invert solve on the box.
* ``sym_eig`` against ``numpy.linalg.eigh``.
* ``from_basis`` / ``spectral_diffusion`` against the float64 torch formula of diffusion_net/layers.py:56-67.
Tolerances are stated per test.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from densematcher_b200 import _lib, spectral_ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mesh(sub, which):
    V0, F = synth.icosphere(sub)
    V = synth.deform(V0, (1.0, 1.3, 0.7)) if which == 1 else synth.deform(V0, (1.2, 0.8, 1.0), bump=0.15, phase=(0.3, 1.1))
    V = V.astype(np.float32).astype(np.float64)          # the goldens' meshes (make_goldens.py: float32 vertices)
    return V, F, synth.cotan_stiffness(V, F), synth.lumped_area(V, F)


def _check_eigenpairs(W, a, evals, Phi, tol_res):
    A = sp.diags(a)
    R = W @ Phi - (A @ Phi) * evals[None]
    res = np.linalg.norm(R / np.sqrt(a)[:, None], axis=0)          # residual of the symmetrised problem
    assert res.max() <= tol_res * max(evals[-1], 1.0), res.max()
    G = Phi.T @ (A @ Phi)
    assert np.abs(G - np.eye(len(evals))).max() < 1e-9
    assert np.all(np.diff(evals) >= -1e-12)


@pytest.mark.parametrize("m", [1, 2, 3, 17, 48, 130, 240])
def test_sym_eig_matches_numpy(m):
    rng = np.random.default_rng(m)
    B = 3
    A = rng.standard_normal((B, m, m))
    A = A + A.transpose(0, 2, 1)
    if m >= 17:                                                      # repeated eigenvalues and a zero block
        Q = np.linalg.qr(rng.standard_normal((m, m)))[0]
        lam = np.concatenate([np.zeros(3), np.ones(4), rng.uniform(1, 50, m - 7)])
        A[1] = (Q * lam) @ Q.T
    w, V = spectral_ops.sym_eig(torch.from_numpy(A).to(DEV))
    w, V = w.cpu().numpy(), V.cpu().numpy()
    for b in range(B):
        wr = np.linalg.eigvalsh(A[b])
        scale = max(np.abs(wr).max(), 1e-300)
        assert np.abs(w[b] - wr).max() <= 1e-12 * scale * max(m, 8)
        assert np.abs(V[b].T @ V[b] - np.eye(m)).max() < 1e-12 * max(m, 8)
        assert np.abs(A[b] @ V[b] - V[b] * w[b][None]).max() <= 1e-12 * scale * max(m, 8)


def test_sym_eig_diagonal_and_identity():
    for A in (np.eye(9), np.diag(np.arange(12.0)[::-1]), np.zeros((5, 5))):
        w, V = spectral_ops.sym_eig(torch.from_numpy(A).to(DEV))
        assert np.allclose(w.cpu().numpy(), np.sort(np.diag(A)), atol=1e-14)
        assert np.allclose(np.abs(V.cpu().numpy().T @ V.cpu().numpy()), np.eye(len(A)), atol=1e-14)


def test_lbo_eigs_matches_reference_spectrum_ico3(golden_fm):
    """642 vertices, k = 20: eigenvalues of the reference's ARPACK run to 1e-8 relative (absolute 1e-8 for the null
    eigenvalue), eigenvectors up to sign to 1e-6."""
    for which in (1, 2):
        V, F, W, a = _mesh(3, which)
        assert np.allclose(a, golden_fm[f"area{which}"], rtol=1e-12)
        evals, Phi, info = spectral_ops.lbo_eigs(W, a, 20, device=DEV, return_info=True)
        assert info["converged"] and info["status"] == 0, info
        evals, Phi = evals.cpu().numpy(), Phi.cpu().numpy()
        ref_ev, ref_Phi = golden_fm[f"evals{which}"], golden_fm[f"Phi{which}"]
        assert np.abs(evals - ref_ev).max() <= 1e-8 * max(ref_ev[-1], 1.0)
        _check_eigenpairs(W, a, evals, Phi, 1e-9)
        cos = np.abs(np.einsum("ik,i,ik->k", Phi, a, ref_Phi))
        assert cos.min() > 1 - 1e-6, cos.min()


def test_lbo_eigs_matches_reference_spectrum_full_size(golden_full):
    """BASELINE size: 2562 vertices, K = 200 (the basis every FM-stage test at full size consumes).  The golden keeps the
    reference's eigenvectors as float32, hence 2e-6 on the vectors; eigenvalues float64, 1e-8 relative."""
    V, F, W, a = _mesh(4, 1)
    evals, Phi, info = spectral_ops.lbo_eigs(W, a, 200, device=DEV, return_info=True)
    assert info["converged"] and info["status"] == 0, info
    evals, Phi = evals.cpu().numpy(), Phi.cpu().numpy()
    assert np.abs(evals - golden_full["evals1"]).max() <= 1e-8 * golden_full["evals1"][-1]
    _check_eigenpairs(W, a, evals, Phi, 1e-9)
    # near-degenerate pairs can mix inside their eigenspace: compare the spectral projectors of clusters instead of single
    # vectors where consecutive eigenvalues are closer than 1e-4 relative
    ref = golden_full["Phi1"]
    ev = golden_full["evals1"]
    M = Phi.T @ (a[:, None] * ref)                               # [gpu, ref] overlaps
    i = 0
    while i < 200:
        j = i + 1
        while j < 200 and ev[j] - ev[j - 1] < 1e-4 * max(ev[j], 1.0):
            j += 1
        if j < 200 or ev[-1] - ev[-2] > 1e-4 * ev[-1]:           # a cluster cut by k = 200 cannot be compared
            s = np.linalg.svd(M[i:j, i:j], compute_uv=False)
            assert s.min() > 1 - 2e-6, (i, j, s.min())
        i = j


def test_lbo_eigs_against_scipy_shift_invert_ragged_sizes():
    """Other sizes / k, against scipy's eigsh(W, k, M=A, sigma=-0.01) run on the box (the reference's call)."""
    for sub, k in ((1, 5), (2, 30), (3, 64)):
        V, F, W, a = _mesh(sub, 2)
        k = min(k, len(a) - 2)
        evals, Phi = spectral_ops.lbo_eigs(W, a, k, device=DEV)
        wr = np.sort(spla.eigsh(W.tocsc(), k=k, M=sp.diags(a).tocsc(), sigma=-0.01)[0])
        assert np.abs(evals.cpu().numpy() - wr).max() <= 1e-8 * max(wr[-1], 1.0)
        _check_eigenpairs(W, a, evals.cpu().numpy(), Phi.cpu().numpy(), 1e-9)


def test_trimesh_process_on_device_matches_host():
    from densematcher_b200.pyFM.mesh import TriMesh
    V, F, _, _ = _mesh(2, 1)
    host = TriMesh(V, F).process(k=25)
    dev = TriMesh(V, F).process(k=25, device=DEV)
    assert np.abs(host.eigenvalues - dev.eigenvalues).max() <= 1e-8 * host.eigenvalues[-1]
    cos = np.abs(np.einsum("ik,i,ik->k", host.eigenvectors, host.vertex_areas, dev.eigenvectors))
    assert cos.min() > 1 - 1e-6


def test_lbo_eigs_many_equals_one_by_one():
    """Several meshes in flight on their own streams / workspaces give the same eigenpairs as sequential calls."""
    meshes = [_mesh(2, 1), _mesh(3, 2), _mesh(2, 2), _mesh(3, 1), _mesh(1, 1)]
    Ws, ms = [m[2] for m in meshes], [m[3] for m in meshes]
    many = spectral_ops.lbo_eigs_many(Ws, ms, 12, device=DEV, n_streams=3)
    for (ev, Phi), W, a in zip(many, Ws, ms):
        ev1, Phi1 = spectral_ops.lbo_eigs(W, a, 12, device=DEV)
        assert torch.equal(ev, ev1) and torch.equal(Phi, Phi1)
        _check_eigenpairs(W, a, ev.cpu().numpy(), Phi.cpu().numpy(), 1e-9)


def test_lbo_eigs_rejects_bad_sizes():
    V, F, W, a = _mesh(1, 1)
    with pytest.raises(ValueError):
        spectral_ops.lbo_eigs(W, a, len(a) + 1, device=DEV)
    with pytest.raises(ValueError):
        spectral_ops.lbo_eigs(W, a, 0, device=DEV)


def test_fps_matches_the_numpy_loop_exactly():
    """Ragged batch of three clouds + a single mesh call, start vertices pinned: identical index sequences."""
    from oracle import dm_oracle as orc
    rng = np.random.default_rng(11)
    sizes = [2000, 777, 1313]
    clouds = [rng.standard_normal((n, 3)) * rng.uniform(0.5, 2.0, 3) for n in sizes]
    clouds[1][5] = clouds[1][9]                                               # duplicate points: ties in the argmax
    off = np.concatenate([[0], np.cumsum(sizes)])
    first = [17, 0, 1312]
    got = spectral_ops.farthest_point_sampling(np.concatenate(clouds), 300, first=first, off=off).cpu().numpy()
    for b, V in enumerate(clouds):
        assert np.array_equal(got[b], orc.fps_euclidean(V, 300, first[b])), b
    V, F, _, _ = _mesh(3, 1)
    from densematcher_b200.pyFM.mesh import TriMesh
    m = TriMesh(V, F)
    assert np.array_equal(m.extract_fps(642, first=3), orc.fps_euclidean(V, 642, 3))      # every vertex, once
    assert len(np.unique(m.extract_fps(100))) == 100                                      # random start
    with pytest.raises(NotImplementedError):
        m.extract_fps(10, geodesic=True)
    with pytest.raises(ValueError):
        spectral_ops.farthest_point_sampling(V, 643)


def test_mesh_zoomout_refine_p2p_subsample_variants(golden_zo):
    """zoomout.py:164-217: explicit subsamples with the vertex map on the samples / on the full meshes, and an integer
    subsample (device FPS) -- against the oracle composition of the same primitives."""
    from oracle import dm_oracle as orc
    from densematcher_b200.pyFM import refine
    from densematcher_b200.pyFM.mesh import TriMesh
    g = golden_zo
    V1, F, _, _ = _mesh(3, 1)
    V2, _, _, _ = _mesh(3, 2)
    m1 = TriMesh.from_basis(g["evals1"], g["Phi1"], g["area1"], V1, F)
    m2 = TriMesh.from_basis(g["evals2"], g["Phi2"], g["area2"], V2, F)
    rng = np.random.default_rng(5)
    p_full = rng.integers(0, 642, 642)
    sub = (g["sub1"], g["sub2"])
    p_sub = rng.integers(0, len(sub[0]), len(sub[1]))
    for p, on_sub in ((p_sub, True), (p_full, False)):
        C, pz = refine.mesh_zoomout_refine_p2p(p, m1, m2, 10, nit=6, step=1, subsample=sub, return_p2p=True,
                                                p2p_on_sub=on_sub)
        Co, po = orc.mesh_zoomout_refine_p2p(p, g["Phi1"], g["Phi2"], g["area2"], 10, nit=6, step=1, subsample=sub,
                                             p2p_on_sub=on_sub, return_p2p=True)
        assert np.abs(C - Co).max() <= 1e-9 * np.abs(Co).max() and np.array_equal(pz, po)
    C = refine.mesh_zoomout_refine_p2p(p_full, m1, m2, 10, nit=4, step=1, subsample=200)
    assert C.shape == (14, 14) and np.all(np.isfinite(C))
    with pytest.raises(ValueError):
        refine.mesh_zoomout_refine_p2p(p_full, m1, m2, 10, nit=4, subsample=200, p2p_on_sub=True)


def _diffusion_reference(x, mass, evals, evecs, t):
    xs = torch.matmul(evecs.transpose(-2, -1), x * mass.unsqueeze(-1))         # geometry.py:572-583
    coefs = torch.exp(-evals.unsqueeze(-1) * t.unsqueeze(0))                   # layers.py:60-61
    return torch.matmul(evecs, coefs * xs)                                     # geometry.py:586-598


def test_from_basis_and_spectral_diffusion_match_the_reference_formula(golden_fm):
    rng = np.random.default_rng(3)
    B, C = 3, 48
    Phi = np.stack([golden_fm["Phi1"], golden_fm["Phi2"], golden_fm["Phi1"][::-1].copy()])
    mass = np.stack([golden_fm["area1"], golden_fm["area2"], golden_fm["area1"][::-1].copy()])
    evals = np.stack([golden_fm["evals1"], golden_fm["evals2"], golden_fm["evals1"]])
    x = rng.standard_normal((B, Phi.shape[1], C)).astype(np.float32)
    t = rng.uniform(0.0, 0.2, C)
    t[0] = 0.0                                                                 # clamped to 1e-8 like layers.py:46-47
    up = lambda a_: torch.from_numpy(np.ascontiguousarray(a_)).to(DEV)
    Phid, md, evd, xd, td = up(Phi), up(mass), up(evals), up(x), up(t)
    ref = _diffusion_reference(xd.double(), md, evd, Phid, torch.clamp(td, min=1e-8))
    coef = rng.standard_normal((B, Phi.shape[2], C))
    fb = spectral_ops.from_basis(up(coef), Phid)
    assert torch.allclose(fb, torch.matmul(Phid, up(coef)), rtol=0, atol=1e-12 * float(np.abs(coef).max()) * 20)
    scale = float(ref.abs().max())
    out64 = spectral_ops.spectral_diffusion(xd, md, evd, Phid, td, flags=_lib.DM_F64_GEMM)
    assert out64.dtype == torch.float32
    assert float((out64.double() - ref).abs().max()) <= 2e-7 * scale           # float32 output rounding only
    out = spectral_ops.spectral_diffusion(xd, md, evd, Phid, td)              # tcgen05 projection: fp32-grade
    assert float((out.double() - ref).abs().max()) <= 1e-5 * scale
    one = spectral_ops.spectral_diffusion(xd[1], md[1], evd[1], Phid[1], td)   # unbatched call
    assert float((one.double() - out[1].double()).abs().max()) <= 1e-6 * scale
    with pytest.raises(ValueError):
        spectral_ops.spectral_diffusion(xd, md, evd, Phid, td[:-1])
