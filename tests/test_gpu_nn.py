"""GPU parity of the fused similarity + argmax path against the CPU oracle and the reference-minted goldens.
Bit-exact index equality is the bar (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import dm_oracle as orc, meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nn():
    from densematcher_b200 import nn as _nn
    return _nn


@pytest.fixture(params=["tcgen05", "ffma"])
def eng(request):
    """Engine flag: the tcgen05 split-bf16 kernel (default) and the CUDA-core fp32 kernel must both be exact."""
    from densematcher_b200 import _lib
    return 0 if request.param == "tcgen05" else _lib.DM_ENGINE_FFMA


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def test_cfg1_matches_reference_kdtree(nn, golden_nn_cfg1, eng):
    """BASELINE config 1: N = M = 2000, d = 384 unit rows; reference = knn_query run in the authoring container."""
    g = golden_nn_cfg1
    F1 = meshgen.random_unit_features(2000, 384, np.random.default_rng(int(g["seed1"])))
    F2 = meshgen.random_unit_features(2000, 384, np.random.default_rng(int(g["seed2"])))
    (p21,), (p12,), stats = nn.nn_argmax(dev(F2), dev(F1), row_epi=(nn.COSINE_UNIT,), col_epi=(nn.COSINE_UNIT,),
                                         return_stats=True, flags=eng)
    assert p21.dtype == torch.int64
    assert np.array_equal(p21.cpu().numpy(), g["ref_p2p_21"])
    assert np.array_equal(p12.cpu().numpy(), g["ref_p2p_12"])
    assert stats[0] < 0.05 * 2000 and stats[1] < 0.05 * 2000      # the float64 path is the exception, not the rule
    # Euclidean form == knn_query, both directions from one pass
    (e21,), (e12,) = nn.nn_argmax(dev(F2), dev(F1), row_epi=(nn.EUCLID,), col_epi=(nn.EUCLID,), flags=eng)
    assert np.array_equal(e21.cpu().numpy(), g["ref_p2p_21"])
    assert np.array_equal(e12.cpu().numpy(), g["ref_p2p_12"])
    # int32 output and the no-recheck fast mode (fp32-grade: allowed to differ on near ties only)
    (q21,), _ = nn.nn_argmax(dev(F2), dev(F1), out_dtype=torch.int32, flags=eng)
    assert q21.dtype == torch.int32 and np.array_equal(q21.cpu().numpy(), g["ref_p2p_21"])
    from densematcher_b200 import _lib
    (r21,), _ = nn.nn_argmax(dev(F2), dev(F1), flags=_lib.DM_NO_RECHECK | eng)
    assert np.count_nonzero(r21.cpu().numpy() != g["ref_p2p_21"]) <= 4


def test_knn_query_shim_small_nonunit(golden_nn_small):
    """Reference-facing knn_query (numpy in/out), non-unit rows, ragged sizes, distances."""
    from densematcher_b200.pyFM.spectral import knn_query
    g = golden_nn_small
    m = knn_query(g["X"], g["Y"])
    assert m.dtype == np.int64 and m.shape == (193,)
    assert np.array_equal(m, g["ref_match"])
    d, m2 = knn_query(g["X"], g["Y"], return_distance=True, n_jobs=4)
    assert np.array_equal(m2, g["ref_match"]) and np.allclose(d, g["ref_dist"], rtol=0, atol=1e-12)
    # float64 inputs take the f64 entry point
    m3 = knn_query(g["X"].astype(np.float64), g["Y"].astype(np.float64))
    assert np.array_equal(m3, g["ref_match"])
    # k > 1: the reference's (n2, k) result, ordered by distance (golden minted from the reference kd-tree)
    d3, m3 = knn_query(g["X"], g["Y"], k=3, return_distance=True)
    assert m3.shape == (193, 3) and m3.dtype == np.int64
    assert np.array_equal(m3, g["ref_match_k3"]) and np.allclose(d3, g["ref_dist_k3"], rtol=0, atol=1e-12)
    assert np.array_equal(knn_query(g["X"], g["Y"], k=3), g["ref_match_k3"])
    assert np.array_equal(knn_query(g["X"], g["Y"], k=16)[:, :3], g["ref_match_k3"])
    with pytest.raises(ValueError):
        knn_query(g["X"][:2], g["Y"], k=3)
    with pytest.raises(ValueError):
        knn_query(g["X"][:0], g["Y"])
    assert knn_query(g["X"], g["Y"][:0]).shape == (0,)


def test_duplicate_rows_resolve_to_lowest_index(nn, eng):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((64, 16)).astype(np.float32)
    X[40] = X[7]
    X[41] = X[7]
    Y = X[[7, 40, 41, 3]]
    (a,), _ = nn.nn_argmax(dev(Y), dev(X), flags=eng)
    assert a.cpu().tolist() == orc.nn_argmax(Y, X).tolist() == [7, 7, 7, 3]
    (e,), _ = nn.nn_argmax(dev(Y), dev(X), row_epi=(nn.EUCLID,), flags=eng)
    assert e.cpu().tolist() == [7, 7, 7, 3]
    # column direction: duplicate query rows
    Yd = np.concatenate([X[:5], X[:5]])
    _, (c,) = nn.nn_argmax(dev(Yd), dev(X), row_epi=(), col_epi=(nn.COSINE_UNIT,), flags=eng)
    assert np.array_equal(c.cpu().numpy(), orc.nn_argmax(Yd, X, axis=0))


@pytest.mark.parametrize("d", [384, 512, 100, 30, 7])
def test_ragged_batch_all_epilogues(nn, d, eng):
    """Ragged batch, every epilogue kind, odd inner dimensions; one launch vs a per-pair oracle loop."""
    rng = np.random.default_rng(100 + d)
    nq = [130, 1, 257, 64, 300]
    nd = [200, 129, 5, 128, 333]
    qo, do = np.concatenate([[0], np.cumsum(nq)]), np.concatenate([[0], np.cumsum(nd)])
    Y = (rng.standard_normal((qo[-1], d)) * rng.uniform(0.3, 2.0, size=(qo[-1], 1))).astype(np.float32)
    X = (rng.standard_normal((do[-1], d)) * rng.uniform(0.3, 2.0, size=(do[-1], 1))).astype(np.float32)
    area = rng.uniform(0.5, 1.5, size=do[-1])
    rbias = rng.standard_normal(qo[-1])
    rows, cols = nn.nn_argmax(dev(Y), dev(X), qo, do,
                              row_epi=(nn.EUCLID, nn.Epi(scale=dev(area))),
                              col_epi=(nn.COSINE, nn.Epi(bias=dev(rbias))), flags=eng)
    rows = [r.cpu().numpy() for r in rows]
    cols = [c.cpu().numpy() for c in cols]
    for p in range(len(nq)):
        y, x = Y[qo[p]:qo[p + 1]], X[do[p]:do[p + 1]]
        assert np.array_equal(rows[0][qo[p]:qo[p + 1]], orc.knn_bruteforce(x, y)), (p, "euclid")
        assert np.array_equal(rows[1][qo[p]:qo[p + 1]], orc.nn_argmax(y, x, col_scale=area[do[p]:do[p + 1]])), (p, "area")
        yn = 1.0 / np.linalg.norm(y.astype(np.float64), axis=1)
        S = (y.astype(np.float64) @ x.astype(np.float64).T)
        assert np.array_equal(cols[0][do[p]:do[p + 1]], (S * yn[:, None]).argmax(0)), (p, "cosine col")
        assert np.array_equal(cols[1][do[p]:do[p + 1]], (S + rbias[qo[p]:qo[p + 1], None]).argmax(0)), (p, "bias col")


def test_recheck_path_is_exact_when_forced(nn, eng):
    """Send EVERY row through the float64 re-evaluation: must equal the oracle on its own."""
    from densematcher_b200 import _lib
    rng = np.random.default_rng(11)
    Y = rng.standard_normal((300, 96)).astype(np.float32)
    X = rng.standard_normal((280, 96)).astype(np.float32)
    rows, cols, st = nn.nn_argmax(dev(Y), dev(X), row_epi=(nn.EUCLID,), col_epi=(nn.EUCLID,),
                                  flags=_lib.DM_RECHECK_ALL | eng, return_stats=True)
    assert st == (300, 280, 580)                                          # every result took the full float64 scan
    assert np.array_equal(rows[0].cpu().numpy(), orc.knn_bruteforce(X, Y))
    assert np.array_equal(cols[0].cpu().numpy(), orc.knn_bruteforce(Y, X))


def test_near_ties_need_the_float64_path(nn, eng):
    """Construct rows whose top-2 gap (1e-9) is far below fp32 resolution: only the recheck can order them."""
    rng = np.random.default_rng(5)
    d = 384
    X = meshgen.random_unit_features(512, d, rng).astype(np.float64)
    Y = meshgen.random_unit_features(64, d, rng).astype(np.float64)
    # make X[2j+1] an almost-copy of X[2j], nudged along y-independent direction so float64 still separates them
    X[1::2] = X[0::2] + 1e-9 * rng.standard_normal((256, d))
    want = orc.nn_argmax(Y, X)
    (got,), _, st = nn.nn_argmax(dev(Y), dev(X), return_stats=True, flags=eng)      # float64 operands
    assert np.array_equal(got.cpu().numpy(), want)
    assert st[0] == 64                                                     # every row was a near tie
    assert st[2] <= 8                                                      # ... decided between two candidates


def test_many_way_near_ties_fall_back_to_the_full_scan(nn, eng):
    """Four almost-identical database rows: the two-candidate shortcut is not enough, the full float64 scan is."""
    rng = np.random.default_rng(6)
    d = 128
    X = meshgen.random_unit_features(256, d, rng).astype(np.float64)
    Y = meshgen.random_unit_features(48, d, rng).astype(np.float64)
    for r in (1, 2, 3):
        X[r::4] = X[0::4] + 1e-10 * rng.standard_normal((64, d))
    want = orc.nn_argmax(Y, X)
    (got,), (gc,), st = nn.nn_argmax(dev(Y), dev(X), col_epi=(nn.COSINE_UNIT,), return_stats=True, flags=eng)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(gc.cpu().numpy(), orc.nn_argmax(Y, X, axis=0))
    assert st[2] >= 48


def test_engine_rounding_error_is_inside_the_bound(nn, eng):
    """The flagging threshold assumes |S~ - S| <= eps |y||x|; eps = (d+4) 2^-24 for the fp32 FMA chain and
    1.2e-5 + (3 ceil(d/64)*4 + 2) 2^-22 for the three split-bf16 tensor-core passes: measure it."""
    rng = np.random.default_rng(9)
    for d in (384, 100, 30):
        Y = (rng.standard_normal((300, d)) * rng.uniform(0.1, 3.0, size=(300, 1))).astype(np.float32)
        X = rng.standard_normal((700, d)).astype(np.float32)
        S = nn.debug_scores(dev(Y), dev(X), flags=eng).cpu().numpy().astype(np.float64)
        S64 = Y.astype(np.float64) @ X.astype(np.float64).T
        kp = (d + 63) // 64 * 64
        eps = (d + 4) * 2.0 ** -24 if eng else 1.2e-5 + (3 * (kp // 16) + 2) * 2.0 ** -22
        bound = eps * np.linalg.norm(Y, axis=1)[:, None] * np.linalg.norm(X, axis=1)[None, :]
        err = np.abs(S - S64)
        assert np.all(err <= bound), (d, float((err / bound).max()))
        assert (err / bound).max() <= 0.5, (d, float((err / bound).max()))   # typical error is far below worst case


def test_empty_and_degenerate_shapes(nn):
    Y = torch.zeros(0, 16, device="cuda")
    X = torch.randn(5, 16, device="cuda")
    (r,), _ = nn.nn_argmax(Y, X)
    assert r.shape == (0,)
    (r,), _ = nn.nn_argmax(torch.randn(3, 16, device="cuda"), X, [0, 0, 3], [0, 2, 5])   # first pair has no queries
    assert r.shape == (3,) and int(r.max()) < 3
    with pytest.raises(ValueError):
        nn.nn_argmax(torch.randn(3, 8, device="cuda"), X)


def test_full_size_properties_cfg3(nn):
    """BASELINE config 3 shape (ragged N ~ U(1500, 2500), d = 384) on a 24-pair sample: properties that do not
    need the oracle -- self-match is the identity, permuting the database permutes the answer, and the
    returned index attains the float64 row maximum."""
    rng = np.random.default_rng(3000)
    P = 24
    nq, nd = rng.integers(1500, 2501, size=P), rng.integers(1500, 2501, size=P)
    qo, do = np.concatenate([[0], np.cumsum(nq)]), np.concatenate([[0], np.cumsum(nd)])
    Y = dev(meshgen.random_unit_features(int(qo[-1]), 384, rng))
    X = dev(meshgen.random_unit_features(int(do[-1]), 384, rng))
    (r,), (c,) = nn.nn_argmax(Y, X, qo, do, col_epi=(nn.COSINE_UNIT,))
    (s,), _ = nn.nn_argmax(X, X, do, do)
    local = torch.cat([torch.arange(n) for n in nd]).cuda()
    assert torch.equal(s, local)
    # float64 check of the attained maximum on the device (torch is only the checker here)
    for p in range(0, P, 5):
        y, x = Y[qo[p]:qo[p + 1]].double(), X[do[p]:do[p + 1]].double()
        S = y @ x.T
        assert torch.equal(S.argmax(1), r[qo[p]:qo[p + 1]])
        assert torch.equal(S.argmax(0), c[do[p]:do[p + 1]])
    # permutation equivariance on one pair
    p = 3
    perm = torch.randperm(int(nd[p]), device="cuda")
    (rp,), _ = nn.nn_argmax(Y[qo[p]:qo[p + 1]], X[do[p]:do[p + 1]][perm])
    assert torch.equal(perm[rp], r[qo[p]:qo[p + 1]])


def test_cfg3_full_size_1024_ragged_pairs(nn):
    """BASELINE config 3 at full size: 1024 pairs, N ~ U(1500, 2500), d = 384, one launch sequence (4 M query rows).
    The batch is generated on the device (torch is plumbing); ALL 1024 pairs are checked against the float64 argmax
    (4.1 M row results + 4.1 M column results), and every returned index is checked to lie inside its own pair."""
    g = torch.Generator(device="cuda").manual_seed(3001)
    rng = np.random.default_rng(3001)
    P = 1024
    nq, nd = rng.integers(1500, 2501, size=P), rng.integers(1500, 2501, size=P)
    qo, do = np.concatenate([[0], np.cumsum(nq)]), np.concatenate([[0], np.cumsum(nd)])
    Y = torch.nn.functional.normalize(torch.randn(int(qo[-1]), 384, device="cuda", generator=g), dim=1)
    X = torch.nn.functional.normalize(torch.randn(int(do[-1]), 384, device="cuda", generator=g), dim=1)
    (r,), (c,), stats = nn.nn_argmax(Y, X, qo, do, col_epi=(nn.COSINE_UNIT,), out_dtype=torch.int32, return_stats=True)
    nd_of_row = torch.repeat_interleave(torch.from_numpy(nd).cuda(), torch.from_numpy(nq).cuda())
    nq_of_col = torch.repeat_interleave(torch.from_numpy(nq).cuda(), torch.from_numpy(nd).cuda())
    assert bool((r >= 0).all()) and bool((r < nd_of_row).all()) and bool((c >= 0).all()) and bool((c < nq_of_col).all())
    assert stats[0] + stats[1] < 0.02 * (qo[-1] + do[-1])
    # EVERY pair against the float64 argmax (torch float64 GEMM on the device as the checker: ~32 MB per pair)
    bad = []
    for p in range(P):
        S = Y[qo[p]:qo[p + 1]].double() @ X[do[p]:do[p + 1]].double().T
        if not (torch.equal(S.argmax(1).int(), r[qo[p]:qo[p + 1]]) and torch.equal(S.argmax(0).int(), c[do[p]:do[p + 1]])):
            bad.append(p)
    assert not bad, bad[:10]


def test_random_shape_sweep_against_float64(nn):
    """Randomised ragged batches (tiny pairs, d from 1 to 530, non-unit rows, every epilogue kind) against the
    float64 argmax; the boundaries of the 128-row / 256-column / 64-K tiling are all crossed."""
    rng = np.random.default_rng(2024)
    for trial in range(25):
        P = int(rng.integers(1, 6))
        nq = rng.integers(1, 700, size=P); nd = rng.integers(1, 700, size=P)
        if trial % 5 == 0:
            nq[0], nd[0] = (128, 256) if trial % 10 == 0 else (129, 257)
        d = int(rng.choice([2, 3, 31, 64, 65, 100, 128, 200, 384, 530]))
        qo, do = np.concatenate([[0], np.cumsum(nq)]), np.concatenate([[0], np.cumsum(nd)])
        Y = (rng.standard_normal((qo[-1], d)) * rng.uniform(0.1, 3)).astype(np.float32)
        X = (rng.standard_normal((do[-1], d)) * rng.uniform(0.1, 3)).astype(np.float32)
        sc = rng.random(do[-1]) + 0.2
        rows = (nn.EUCLID, nn.Epi(scale=dev(sc)))
        cols = (nn.COSINE, nn.COSINE_UNIT)
        (e, w), (cc, cu) = nn.nn_argmax(dev(Y), dev(X), qo, do, row_epi=rows, col_epi=cols)
        for p in range(P):
            y, x = Y[qo[p]:qo[p + 1]].astype(np.float64), X[do[p]:do[p + 1]].astype(np.float64)
            S = y @ x.T
            tag = (trial, p, int(nq[p]), int(nd[p]), d)
            assert np.array_equal(e[qo[p]:qo[p + 1]].cpu().numpy(), (S - 0.5 * (x * x).sum(1)[None]).argmax(1)), tag
            assert np.array_equal(w[qo[p]:qo[p + 1]].cpu().numpy(), (S * sc[do[p]:do[p + 1]][None]).argmax(1)), tag
            if d >= 3:  # in 1-2 dimensions cosines tie exactly up to rounding: the float64 winner is rounding noise
                assert np.array_equal(cc[do[p]:do[p + 1]].cpu().numpy(),
                                      (S * (1.0 / np.linalg.norm(y, axis=1))[:, None]).argmax(0)), tag
            assert np.array_equal(cu[do[p]:do[p + 1]].cpu().numpy(), S.argmax(0)), tag
