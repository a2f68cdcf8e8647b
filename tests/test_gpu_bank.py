"""Mesh bank (dm_bank_prepare / dm_match_bank_pairs): the once-per-mesh preparation + pairs as id lists must give, bit for
bit, what the per-pair path (dm_match_pairs on the assembled batch) gives -- itself pinned on the oracle and on the
reference's goldens by the other GPU tests -- and a sample is checked against the oracle directly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = ("nn_p2p_21", "nn_p2p_12", "C", "p2p_21", "p2p_12", "p2p_21_adjoint", "p2p_12_adjoint")


def _bank(sizes, d, K, seed, dup=True):
    import torch  # noqa: F401
    from densematcher_b200 import pipeline, synth
    rng = np.random.default_rng(seed)
    bases = [synth.synthetic_basis(int(n), K, rng) for n in sizes]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    F = synth.random_unit_features(int(off[-1]), d, rng)
    if dup:  # exact duplicates inside a mesh: ties that only the float64 re-evaluation orders (lowest index)
        F[off[1] + 77] = F[off[1] + 3]
        F[off[0] + 5] = F[off[0] + 120]
    bank = pipeline.MeshBankDevice(F, off, Phi=np.concatenate([b[1] for b in bases]), evals=np.stack([b[0] for b in bases]),
                                   area=np.concatenate([b[2] for b in bases]))
    return bank, bases, F, off


def _same(a, b):
    for n in NAMES:
        x, y = a[n].cpu().numpy(), b[n].cpu().numpy()
        assert x.dtype == y.dtype and np.array_equal(x, y), n


@pytest.mark.parametrize("sizes,d,K,k", [
    ([512] * 5, 384, 100, 100),               # uniform, even row tiles, long contraction: the CTA-pair score kernel
    ([300, 517, 128, 255, 641, 65], 384, 40, 40),  # ragged: tails in every tile position
    ([200, 333, 260, 129], 100, 64, 30),      # k < K (leading dimension > k), short contractions
])
def test_bank_pairs_equal_assembled_pairs(sizes, d, K, k):
    import torch
    from densematcher_b200 import pipeline
    bank, _, _, _ = _bank(sizes, d, K, seed=11 + len(sizes))
    M = len(sizes)
    src = np.array([i for i in range(M) for j in range(M) if i != j] + [0, 1], dtype=np.int64)
    dst = np.array([j for i in range(M) for j in range(M) if i != j] + [0, 1], dtype=np.int64)  # + a mesh against itself
    assert bank.bank_supported(k)
    got = {}
    for dt in (torch.int64, torch.int32):
        got = bank.match(src, dst, k=k, out_dtype=dt)
        ref = pipeline.match_pairs_device(bank.assemble(src, dst), k=k, out_dtype=dt)
        _same(got, ref)
    # the chunked driver (prepared state reused, several calls) = one call
    chunks, _ = pipeline.match_bank_pairs(bank, src, dst, chunk_pairs=7, k=k, to_host=False)
    for n in NAMES:
        assert np.array_equal(torch.cat([c[n] for c in chunks]).cpu().numpy(), got[n].cpu().numpy()), n


def test_bank_pairs_against_the_oracle():
    from oracle import dm_oracle as orc
    bank, bases, F, off = _bank([400, 380, 420], 64, 24, seed=3, dup=False)
    src, dst = np.array([0, 2, 1]), np.array([1, 0, 2])
    k = 20
    got = bank.match(src, dst, k=k, w_descr=1e4, w_lap=1e3)
    o1, o2 = got["off1"], got["off2"]
    for p, (i, j) in enumerate(zip(src, dst)):
        F1, F2 = F[off[i]:off[i + 1]], F[off[j]:off[j + 1]]
        assert np.array_equal(got["nn_p2p_21"][o2[p]:o2[p + 1]].cpu().numpy(), orc.nn_argmax(F2, F1))
        assert np.array_equal(got["nn_p2p_12"][o1[p]:o1[p + 1]].cpu().numpy(), orc.nn_argmax(F1, F2))
        (ev1, P1, a1), (ev2, P2, a2) = bases[i], bases[j]
        C = got["C"][p].cpu().numpy()
        Co = orc.fmap_solve_closed_form(orc.project(P1, a1, F1, k), orc.project(P2, a2, F2, k), ev1[:k], ev2[:k],
                                        orc.fmap_c00(P1, P2, a1, a2), 1e4, 1e3)
        assert np.linalg.norm(C - Co) <= 1e-4 * np.linalg.norm(Co)   # north-star bar for C (measured ~1e-6)
        r21, r12, MI = orc.fm_to_p2p(C, P1[:, :k], P2[:, :k], a1)
        assert np.array_equal(got["p2p_21_adjoint"][o2[p]:o2[p + 1]].cpu().numpy(), r21)
        assert np.array_equal(got["p2p_12_adjoint"][o1[p]:o1[p + 1]].cpu().numpy(), r12)
        assert np.array_equal(got["p2p_21"][o2[p]:o2[p + 1]].cpu().numpy(), MI.argmax(1))
        assert np.array_equal(got["p2p_12"][o1[p]:o1[p + 1]].cpu().numpy(), MI.argmax(0))


def test_bank_host_entry_and_bad_ids():
    import torch
    from densematcher_b200 import fm, pipeline
    bank, bases, F, off = _bank([256, 256, 300, 190], 48, 16, seed=9)
    src, dst = pipeline.intra_category_pairs(np.array([0, 0, 1, 1]))
    k = 12
    dev = bank.match(src, dst, k=k)
    hbank = pipeline.MeshBankHost(F=F, off=off, Phi=np.concatenate([b[1] for b in bases]),
                                  evals=np.stack([b[0] for b in bases]), area=np.concatenate([b[2] for b in bases]))
    for _ in range(2):  # the second call reuses the stager's buffers (and re-prepares the freshly uploaded bank)
        host = pipeline.match_bank_pairs_host(hbank, src, dst, chunk_pairs=3, k=k)
        for n in NAMES:
            assert np.array_equal(host[n], dev[n].cpu().numpy()), n
    # float32 eigenvectors (the operator cache's dtype) are widened on the device: same as float64 input of those values
    h32 = pipeline.MeshBankHost(F=F, off=off, Phi=hbank.Phi.astype(np.float32), evals=hbank.evals, area=hbank.area)
    b32 = pipeline.MeshBankDevice(F, off, Phi=hbank.Phi.astype(np.float32).astype(np.float64), evals=hbank.evals, area=hbank.area)
    host = pipeline.match_bank_pairs_host(h32, src, dst, chunk_pairs=4, k=k)
    ref = b32.match(src, dst, k=k)
    for n in NAMES:
        assert np.array_equal(host[n], ref[n].cpu().numpy()), n
    # ids out of range / offsets that do not match the meshes are reported, not read out of bounds
    st = bank.prepared(k)
    ids = torch.tensor([0, 7], dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        fm.match_bank_pairs(st, ids, ids, np.array([0, 256, 512]), np.array([0, 256, 512]), 1e4, 1e3)
    ids = torch.tensor([0, 2], dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        fm.match_bank_pairs(st, ids, ids, np.array([0, 256, 512]), np.array([0, 256, 512]), 1e4, 1e3)


def test_bank_prepared_piecewise():
    """A rank of a sharded job prepares only the meshes its block of pairs touches; later pairs extend the range, each
    mesh is prepared once, and the results do not depend on how the bank was prepared."""
    import torch
    from densematcher_b200 import fm, pipeline
    bank, _, _, _ = _bank([260, 300, 256, 384, 200, 310], 96, 32, seed=21)
    k = 32
    full = pipeline.MeshBankDevice(bank.F, bank.off_h, Phi=bank.Phi, evals=bank.evals, area=bank.area)
    src, dst = np.array([2, 3, 3, 2]), np.array([3, 2, 3, 2])
    st = bank.prepared(k, (2, 4))
    assert (st.lo, st.hi) == (2, 4)
    _same(bank.match(src, dst, k=k), full.match(src, dst, k=k))
    assert (st.lo, st.hi) == (2, 4)
    # ids outside the prepared range are refused by the library call itself
    ids = torch.tensor([0, 2], dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        fm.match_bank_pairs(st, ids, ids, np.array([0, 260, 516]), np.array([0, 260, 516]), 1e4, 1e3)
    # pairs that reach further extend the prepared range on both sides
    src2, dst2 = np.array([0, 5, 1, 4]), np.array([5, 0, 4, 1])
    got = bank.match(src2, dst2, k=k)
    assert (st.lo, st.hi) == (0, 6)
    _same(got, full.match(src2, dst2, k=k))
    _same(got, pipeline.match_pairs_device(bank.assemble(src2, dst2), k=k))
    # the sharded driver prepares the range of its block
    b2 = pipeline.MeshBankDevice(bank.F, bank.off_h, Phi=bank.Phi, evals=bank.evals, area=bank.area)
    s_all, d_all = pipeline.intra_category_pairs(np.array([0, 0, 0, 1, 1, 1]))
    chunks, (lo, hi) = pipeline.match_bank_pairs(b2, s_all, d_all, chunk_pairs=4, rank=1, world=2, k=k, to_host=False)
    assert (b2._states[k].lo, b2._states[k].hi) == (3, 6) and (lo, hi) == (6, 12)
    ref = full.match(s_all[lo:hi], d_all[lo:hi], k=k)
    for n in NAMES:
        assert np.array_equal(torch.cat([c[n] for c in chunks]).cpu().numpy(), ref[n].cpu().numpy()), n
