"""GPU linear sum assignment (dm_lap_solve) against scipy.optimize.linear_sum_assignment -- the solver the reference
itself calls (functional_map.py:57,66,78).  Index outputs must be IDENTICAL to scipy's, ties included."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from oracle import dm_oracle as orc, meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fm():
    from densematcher_b200 import fm as _fm
    return _fm


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def check(fm, mats, maximize):
    got = fm.lap_solve([dev(m) for m in mats], maximize=maximize)
    for m, (r, c) in zip(mats, got):
        rr, cc = linear_sum_assignment(m, maximize=maximize)
        assert r.dtype == np.int64 and c.dtype == np.int64
        assert np.array_equal(r, rr) and np.array_equal(c, cc)


@pytest.mark.parametrize("maximize", [False, True])
def test_random_float_shapes(fm, maximize):
    rng = np.random.default_rng(1)
    shapes = [(1, 1), (1, 7), (7, 1), (50, 50), (40, 60), (60, 40), (200, 200), (513, 700), (700, 513), (1025, 1030)]
    check(fm, [rng.standard_normal(s) for s in shapes], maximize)


@pytest.mark.parametrize("maximize", [False, True])
def test_ties_follow_scipy(fm, maximize):
    rng = np.random.default_rng(2)
    mats = [rng.integers(0, 4, (30, 30)), rng.integers(0, 4, (30, 45)), rng.integers(0, 4, (45, 30)),
            rng.integers(0, 2, (64, 64)), np.ones((20, 20)), np.zeros((17, 33)), np.zeros((33, 17)),
            rng.integers(0, 3, (300, 300)), np.round(rng.standard_normal((128, 256)), 1)]
    check(fm, [np.asarray(m, dtype=np.float64) for m in mats], maximize)


def test_single_tensor_and_inf_entries(fm):
    rng = np.random.default_rng(3)
    m = rng.standard_normal((40, 40))
    m[rng.random((40, 40)) < 0.3] = np.inf   # forbidden edges are allowed as long as a matching exists
    np.fill_diagonal(m, rng.standard_normal(40))
    r, c = fm.lap_solve(dev(m))
    rr, cc = linear_sum_assignment(m)
    assert np.array_equal(r, rr) and np.array_equal(c, cc)


def test_invalid_and_infeasible_raise_like_scipy(fm):
    m = np.ones((5, 5))
    m[2, 3] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        fm.lap_solve(dev(m))
    m[2, 3] = -np.inf
    with pytest.raises(ValueError, match="invalid numeric"):
        fm.lap_solve(dev(m))
    m[2, 3] = np.inf
    with pytest.raises(ValueError, match="invalid numeric"):
        fm.lap_solve(dev(m), maximize=True)
    m = np.full((4, 4), np.inf)
    m[:, 0] = 1.0
    with pytest.raises(ValueError, match="infeasible"):
        fm.lap_solve(dev(m))
    # per-problem status instead of raising
    res, st = fm.lap_solve([dev(np.eye(3)), dev(m)], return_status=True)
    assert st.tolist() == [0, 2] and np.array_equal(res[0][1], linear_sum_assignment(np.eye(3))[1])


def test_empty_problems(fm):
    res = fm.lap_solve([dev(np.zeros((0, 5))), dev(np.zeros((4, 0))), dev(np.eye(3))])
    assert [len(r) for r, _ in res] == [0, 0, 3]


def _mapped_indicator(sub, k, seed):
    V, F = meshgen.icosphere(sub)
    V1 = meshgen.deform(V, (1.0, 0.9, 1.1), 0.1, (0.3, 0.2))
    V2 = meshgen.deform(V, (1.1, 1.0, 0.85), 0.15, (1.0, 0.5))
    ev1, P1, a1 = meshgen.lbo_basis(V1, F, k)
    ev2, P2, a2 = meshgen.lbo_basis(V2, F, k)
    rng = np.random.default_rng(seed)
    F1 = meshgen.bandlimited_features(P1, 64, k, rng, dtype=np.float64)
    F2 = F1 + 0.05 * rng.standard_normal(F1.shape)
    C = orc.fmap_solve_closed_form(orc.project(P1, a1, F1), orc.project(P2, a2, F2), ev1, ev2,
                                   orc.fmap_c00(P1, P2, a1, a2), 1e4, 1e3)
    return (P2 @ C @ P1.T) * a1[None, :]


def test_mapped_indicator_assignment_matches_scipy(fm):
    """The reference's actual call: maximise the dense N2 x N1 mapped indicator (functional_map.py:57)."""
    mi3 = _mapped_indicator(3, 30, 0)     # 642 x 642
    mi4 = _mapped_indicator(4, 50, 1)     # 2562 x 2562, ~55k Dijkstra steps
    eta = np.ones(mi4.shape[0])
    cost4 = mi4 * eta[:, None] - 1000 * (1 - eta[:, None])
    check(fm, [mi3, cost4, mi3[:600], mi3[:, :600]], True)
