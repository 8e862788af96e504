"""GPU tests of the device BPR sampler (csrc/sampler.cu) against the contract of util/sampler.py:237-264:
every training pair appears once per epoch, negatives are never training items of the user and are uniform over
the rest, batches have the reference's shapes (short last batch), streams are reproducible."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from hypergraph_diffusion_for_recommendation_b200 import _lib, sampler

    assert torch.cuda.is_available()
    _lib.lib()
    return sampler


def graph(n_users=300, n_items=500, n_train=20_000, seed=4):
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(n_users, n_items, n_train, seed=seed)
    return g.train_u, g.train_i


def test_epoch_covers_every_pair_once_and_negatives_avoid_the_training_set(S):
    u, i = graph()
    s = S.PairwiseSampler(u, i, 300, 500, seed=7)
    train = set(zip(u.tolist(), i.tolist()))
    seen, sizes = [], []
    for bu, bp, bn in s.epoch(4096):
        assert bu.dtype == torch.int64 and bu.is_cuda and bu.shape == bp.shape == bn.shape
        sizes.append(bu.numel())
        bu, bp, bn = bu.cpu().numpy(), bp.cpu().numpy(), bn.cpu().numpy()
        seen += list(zip(bu.tolist(), bp.tolist()))
        assert all((a, c) not in train for a, c in zip(bu.tolist(), bn.tolist()))
        assert bn.min() >= 0 and bn.max() < 500
    assert sizes == [4096] * 4 + [20_000 - 4 * 4096]          # the reference's short last batch
    assert sorted(seen) == sorted(train) and len(seen) == len(train)  # a permutation of the training pairs
    assert int(s.gave_up.item()) == 0


def test_negatives_are_uniform_over_the_non_training_items(S):
    # one user who interacted with items 0..99 of 200: negatives must be uniform over 100..199
    u = np.zeros(100, dtype=np.int64)
    i = np.arange(100, dtype=np.int64)
    s = S.PairwiseSampler(u, i, 1, 200, seed=1)
    counts = np.zeros(200, dtype=np.int64)
    for _ in range(200):
        for _, _, bn in s.epoch(100, n_negs=5):
            counts += np.bincount(bn.cpu().numpy(), minlength=200)
    assert counts[:100].sum() == 0
    expect = counts.sum() / 100
    chi2 = ((counts[100:] - expect) ** 2 / expect).sum()
    assert chi2 < 160, chi2  # 99 degrees of freedom: P(chi2 > 160) ~ 1e-4


def test_streams_are_reproducible_and_layout_of_multiple_negatives(S):
    u, i = graph(seed=9)
    a = [tuple(t.cpu() for t in b) for b in S.PairwiseSampler(u, i, 300, 500, seed=3).epoch(5000, n_negs=3)]
    b = [tuple(t.cpu() for t in b) for b in S.PairwiseSampler(u, i, 300, 500, seed=3).epoch(5000, n_negs=3)]
    c = [tuple(t.cpu() for t in b) for b in S.PairwiseSampler(u, i, 300, 500, seed=4).epoch(5000, n_negs=3)]
    assert all(torch.equal(x, y) for ba, bb in zip(a, b) for x, y in zip(ba, bb))
    assert not all(torch.equal(ba[2], bc[2]) for ba, bc in zip(a, c))
    assert a[0][2].numel() == 3 * a[0][0].numel()  # j_idx holds n_negs negatives per positive, positive-major
    # reference-signature generator on an Interaction-like object
    data = type("D", (), {})()
    data.training_data = [[int(x) + 1000, int(y) + 5000, 1.0] for x, y in zip(u[:500], i[:500])]
    data.user = {int(x) + 1000: int(x) for x in np.unique(u)}
    data.item = {int(y) + 5000: int(y) for y in np.unique(i)}
    data.n_users, data.n_items = 300, 500
    got = list(S.next_batch_pairwise(data, 128))
    assert len(got) == 4 and got[-1][0].numel() == 500 - 3 * 128


def test_dense_user_gives_up_gracefully(S):
    # a user who has every item: no valid negative exists; the kernel must terminate and report it
    u = np.zeros(50, dtype=np.int64)
    i = np.arange(50, dtype=np.int64)
    s = S.PairwiseSampler(u, i, 1, 50, seed=2)
    batches = list(s.epoch(50))
    assert batches[0][2].numel() == 50 and int(s.gave_up.item()) == 50
