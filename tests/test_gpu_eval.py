"""GPU parity of full-ranking evaluation (csrc/eval_topk.cu through the C ABI) against the CPU oracle and
the reference's golden vectors.  Bar: top-K item ids AND scores bit-exact (ties by ascending item id)."""
import numpy as np
import pytest

from oracle import hgr_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import torch

    from hypergraph_diffusion_for_recommendation_b200 import _lib, evaluation

    assert torch.cuda.is_available()
    _lib.lib()
    return evaluation


def cuda(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def train_csr(u, i, n_users):
    order = np.lexsort((i, u))
    u, i = np.asarray(u)[order], np.asarray(i)[order]
    keep = np.ones(u.size, bool)
    keep[1:] = (u[1:] != u[:-1]) | (i[1:] != i[:-1])
    u, i = u[keep], i[keep]
    ptr = np.zeros(n_users + 1, np.int64)
    np.cumsum(np.bincount(u, minlength=n_users), out=ptr[1:])
    return ptr, i.astype(np.int32)


def run(E, ue, ie, test_users, ptr, idx, k, mode, engine):
    ids, sc, stats = E.fullrank_topk(cuda(ue), cuda(ie), cuda(np.asarray(test_users, np.int32)), cuda(ptr), cuda(idx), k, mode=mode,
                                     engine=engine, return_stats=True)
    return ids.cpu().numpy().astype(np.int64), sc.cpu().numpy(), stats.cpu().numpy()


def check_exact(E, ue, ie, test_users, ptr, idx, k, engines=("simt", "tensor"), modes=("exact", "refquirk")):
    out = {}
    for mode in modes:
        want_ids, want_sc = O.fullrank_topk(ue, ie, test_users, ptr, idx, k, mode=mode)
        for engine in engines:
            ids, sc, stats = run(E, ue, ie, test_users, ptr, idx, k, mode, engine)
            bad = np.nonzero((ids != want_ids).any(axis=1))[0]
            assert bad.size == 0, "%s/%s: %d users differ, first %d: got %s want %s" % (
                mode, engine, bad.size, bad[0], ids[bad[0]], want_ids[bad[0]])
            assert np.array_equal(sc.view(np.uint32), want_sc.view(np.uint32)), "%s/%s scores not bit-exact" % (mode, engine)
            assert stats[3] < 1_000_000, "tensor-score error reached %.2f of the proven bound" % (stats[3] / 1e6)
            out[(mode, engine)] = stats
    return out


def test_golden_reference_rec_lists_and_metric_strings(E, golden, pl_graph):
    """The reference's own rec lists (duplicate quirk included) and its Recall/NDCG strings."""
    id2item = golden["pl_id2item"]
    user = {r: k for k, r in enumerate(golden["pl_id2user"])}
    item = {int(r): k for k, r in enumerate(id2item)}
    users_raw = golden["eval_users_raw"]
    test_users = np.array([user[int(r)] for r in users_raw])
    tip, tix, _ = O.interaction_matrix(pl_graph["u"], pl_graph["i"], pl_graph["n_users"], pl_graph["n_items"])
    for engine in ("tensor", "simt"):
        ids, sc, _ = run(E, golden["eval_user_emb"], golden["eval_item_emb"], test_users, tip, tix.astype(np.int32), 20, "refquirk", engine)
        assert np.array_equal(id2item[ids], golden["eval_rec_items_raw"]), engine
        assert np.abs(sc - golden["eval_rec_scores"]).max() <= 1e-5 * np.abs(golden["eval_rec_scores"]).max()
        truth = {int(r): [] for r in users_raw}
        for uu, ii, _ in golden["pl_test"]:
            if int(uu) in truth:
                truth[int(uu)].append(int(ii))
        ptr = np.zeros(len(users_raw) + 1, np.int64)
        np.cumsum([len(truth[int(r)]) for r in users_raw], out=ptr[1:])
        items = np.array([item.get(x, -1) for r in users_raw for x in truth[int(r)]], dtype=np.int64)
        strings = E.ranking_evaluation_ids(ptr, items, ids, [10, 20])
        assert strings == [str(s) for s in golden["eval_measures"]], engine
        # the same strings with the per-user work on the device (hgr_rank_metrics)
        assert E.ranking_evaluation_device(ptr, items, cuda(ids.astype(np.int32)), [10, 20]) == strings
    check_exact(E, golden["eval_user_emb"], golden["eval_item_emb"], test_users, tip, tix.astype(np.int32), 20)


@pytest.mark.parametrize("k", [1, 10, 20, 40, 64])
def test_random_and_trained_like_embeddings(E, k):
    rng = np.random.default_rng(k)
    n_users, n_items, d = 700, 3000, 64
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(n_users, n_items, 40_000, seed=3)
    ptr, idx = train_csr(g.train_u, g.train_i, n_users)
    ue = (rng.standard_normal((n_users, d)) * 0.1).astype(np.float32)
    ie = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    test_users = rng.permutation(n_users)[:600]
    check_exact(E, ue, ie, test_users, ptr, idx, k)
    # "trained": the user's training items get the highest scores, so the mask decides the list
    ue2 = ue.copy()
    for u in range(n_users):
        its = idx[ptr[u]:ptr[u + 1]]
        if its.size:
            ue2[u] = ie[its].mean(axis=0) * 3
    ie2 = ie * (1 + 4 * rng.random((n_items, 1))).astype(np.float32)  # uneven item norms
    stats = check_exact(E, ue2, ie2, test_users, ptr, idx, k)
    s = stats[("exact", "tensor")]
    assert s[2] == 0, "no user should need the brute-force fallback here (%d did)" % s[2]
    # LayerNorm-like tables: a large common component, so all scores of a user cluster within a fraction of
    # a percent of ||u|| ||i|| -- the split-bf16 error bound must still separate them
    ue3 = (1.0 + 0.05 * rng.standard_normal((n_users, d))).astype(np.float32)
    ie3 = (1.0 + 0.05 * rng.standard_normal((n_items, d))).astype(np.float32)
    stats = check_exact(E, ue3, ie3, test_users, ptr, idx, k, modes=("exact",))
    assert stats[("exact", "tensor")][2] == 0


def test_ties_zero_and_duplicate_rows(E):
    rng = np.random.default_rng(5)
    n_users, n_items, d, k = 40, 500, 64, 20
    ptr, idx = train_csr(rng.integers(0, n_users, 600), rng.integers(0, n_items, 600), n_users)
    test_users = np.arange(n_users)
    # all-zero embeddings: every score ties, the list is the K lowest unmasked ids
    z = np.zeros((n_users, d), np.float32)
    check_exact(E, z, np.zeros((n_items, d), np.float32), test_users, ptr, idx, k, modes=("exact",))
    # refquirk with exact ties among the first K candidates: numba's sort is stable only up to 15 elements
    check_exact(E, z, np.zeros((n_items, d), np.float32), test_users, ptr, idx, 10)
    # many identical item rows: exact score ties broken by item id
    ie = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    ie[100:200] = ie[7]
    ie[300:340] = ie[3]
    ue = (rng.standard_normal((n_users, d)) * 0.1).astype(np.float32)
    check_exact(E, ue, ie, test_users, ptr, idx, k, modes=("exact",))
    check_exact(E, ue, ie, test_users, ptr, idx, 12)


def test_user_with_fewer_unmasked_items_than_k(E):
    rng = np.random.default_rng(6)
    n_users, n_items, d, k = 6, 130, 64, 20
    tu, ti = [], []
    for u, n_train in enumerate([125, 130, 111, 0, 1, 129]):  # 5, 0, 19, 130, 129, 1 unmasked items
        its = rng.permutation(n_items)[:n_train]
        tu += [u] * n_train
        ti += list(its)
    ptr, idx = train_csr(np.array(tu, np.int64), np.array(ti, np.int64), n_users)
    ue = (rng.standard_normal((n_users, d)) * 0.1).astype(np.float32)
    ie = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    check_exact(E, ue, ie, np.arange(n_users), ptr, idx, k, modes=("exact",))


def test_item_range_splits_and_ragged_sizes(E):
    """Few users x many items: the item range is split across CTAs; sizes not multiples of the tiles."""
    rng = np.random.default_rng(7)
    n_users, n_items, d, k = 300, 20_011, 64, 20
    ptr, idx = train_csr(rng.integers(0, n_users, 9000), rng.integers(0, n_items, 9000), n_users)
    ue = (rng.standard_normal((n_users, d)) * 0.1).astype(np.float32)
    ie = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    test_users = rng.permutation(n_users)[:257]
    stats = check_exact(E, ue, ie, test_users, ptr, idx, k, modes=("exact",))
    s = stats[("exact", "tensor")]
    assert s[0] < 257 * 4000 and s[1] <= s[0]  # far fewer candidates than scores (5.1 M)


def test_candidate_overflow_falls_back_to_exact_path(E):
    """Scores increasing with the item id make every item a new best: the candidate list overflows and the
    brute-force kernel takes the user over.  The result must not change."""
    n_users, n_items, d, k = 5, 6000, 64, 20
    ue = np.zeros((n_users, d), np.float32)
    ue[:, 0] = 1.0
    ie = np.zeros((n_items, d), np.float32)
    ie[:, 0] = np.linspace(0.1, 5.0, n_items, dtype=np.float32)
    ptr, idx = train_csr(np.array([0, 0, 1]), np.array([5999, 5998, 17]), n_users)
    stats = check_exact(E, ue, ie, np.arange(n_users), ptr, idx, k, modes=("exact",))
    assert stats[("exact", "tensor")][2] == n_users


@pytest.mark.parametrize("d", [32, 128])
def test_other_widths_use_the_simt_engine(E, d):
    rng = np.random.default_rng(d)
    n_users, n_items, k = 50, 900, 20
    ptr, idx = train_csr(rng.integers(0, n_users, 700), rng.integers(0, n_items, 700), n_users)
    ue = (rng.standard_normal((n_users, d)) * 0.1).astype(np.float32)
    ie = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    check_exact(E, ue, ie, np.arange(n_users), ptr, idx, k, engines=("auto",))
    from hypergraph_diffusion_for_recommendation_b200 import _lib

    with pytest.raises(_lib.HgrError):
        run(E, ue, ie, np.arange(n_users), ptr, idx, k, "exact", "tensor")


def test_engines_agree_at_gowalla_scale(E):
    """Size-independent property at a BASELINE shape (30 k users x 41 k items): the tensor path and the fp32
    brute force return identical ids and scores; ids are unique, unmasked and sorted by score."""
    import torch

    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    n_users, n_items, d, k = 30_000, 41_000, 64, 20
    g = powerlaw_interactions(n_users, n_items, 1_000_000, seed=1234)
    ev = E.EvalData.from_arrays(n_users, n_items, g.train_u, g.train_i, g.test_u, g.test_i)
    torch.manual_seed(0)
    ue = torch.nn.init.xavier_uniform_(torch.empty(n_users, d)).cuda()
    ie = torch.nn.init.xavier_uniform_(torch.empty(n_items, d)).cuda()
    sub = ev.test_users[:4096]
    a_ids, a_sc, stats = E.fullrank_topk(ue, ie, ev.test_users, ev.train_indptr, ev.train_indices, k, engine="tensor", return_stats=True)
    b_ids, b_sc = E.fullrank_topk(ue, ie, sub, ev.train_indptr, ev.train_indices, k, engine="simt")
    assert torch.equal(a_ids[:4096], b_ids) and torch.equal(a_sc[:4096], b_sc)
    assert int(stats[2]) == 0 and int(stats[3]) < 1_000_000
    ids = a_ids.cpu().numpy()
    sc = a_sc.cpu().numpy()
    assert (np.diff(sc, axis=1) <= 0).all()
    assert (np.sort(ids, axis=1)[:, 1:] != np.sort(ids, axis=1)[:, :-1]).all()
    ptr, idx = ev.train_indptr.cpu().numpy(), ev.train_indices.cpu().numpy()
    users = ev.test_users.cpu().numpy()
    for r in range(0, users.size, 997):
        assert not np.isin(ids[r], idx[ptr[users[r]]:ptr[users[r] + 1]]).any()


def test_device_metric_reducer_matches_host_and_oracle_with_duplicates_and_unknown_items(E):
    rng = np.random.default_rng(21)
    n_users, n_items, k = 4000, 3000, 40
    truth = [rng.choice(n_items, rng.integers(1, 60), replace=False).astype(np.int64) for _ in range(n_users)]
    for t in truth[::7]:
        t[0] = -1                                    # an item never seen in training: counts in |truth|, never hit
    rec = rng.integers(0, n_items, (n_users, k))
    rec[:, 9] = rec[:, 3]                            # duplicated recommendations (the reference's quirk)
    rec[::5, 0] = [t[-1] if t[-1] >= 0 else 0 for t in truth[::5]]  # guaranteed hits (never the -1 placeholder)
    ptr = np.zeros(n_users + 1, np.int64)
    np.cumsum([len(t) for t in truth], out=ptr[1:])
    items = np.concatenate(truth)
    for top in ([10, 20, 40], [20], [5, 40, 50]):
        host = E.ranking_evaluation_ids(ptr, items, rec, top)
        assert E.ranking_evaluation_device(ptr, items, cuda(rec.astype(np.int32)), top) == host
    assert host[:1] == ["Top 5\n"]
    want = O.ranking_evaluation([list(t) for t in truth], rec, [10, 20, 40])
    assert E.ranking_evaluation_device(ptr, items, cuda(rec.astype(np.int32)), [10, 20, 40]) == want


def test_device_metric_sums_reproduce_the_python_sums_bit_for_bit(E):
    """hgr_rank_metric_sums against the python loops it replaces (host_sums=True), on 200 000 users: the sums are sequential
    double additions in user order (Neumaier-compensated for the recall list under python >= 3.12), so the UNROUNDED values are
    equal, not just the 5-decimal strings."""
    import sys

    import torch

    rng = np.random.default_rng(33)
    n_users, n_items, k = 200_000, 5000, 20
    n_truth = rng.integers(1, 40, n_users)
    ptr = np.zeros(n_users + 1, np.int64)
    np.cumsum(n_truth, out=ptr[1:])
    items = rng.integers(0, n_items, int(ptr[-1]))
    rec = rng.integers(0, n_items, (n_users, k)).astype(np.int32)
    rec[::3, 1] = items[ptr[:-1][::3]]  # plenty of hits
    rd = cuda(rec)
    for top in ([10, 20], [20]):
        assert E.ranking_evaluation_device(ptr, items, rd, top) == E.ranking_evaluation_device(ptr, items, rd, top, host_sums=True)
    # the raw sums of one call, device against numpy / python arithmetic
    import ctypes as C
    import math

    from hypergraph_diffusion_for_recommendation_b200 import _lib

    hits_ref = np.array([len(set(items[ptr[r]:ptr[r + 1]].tolist()) & set(rec[r].tolist())) for r in range(0, n_users, 997)])
    lib = _lib.lib()
    tp = cuda(ptr)
    order = np.lexsort((items, np.repeat(np.arange(n_users), n_truth)))
    ti = cuda(items[order].astype(np.int32))
    top_dev = cuda(np.array([20], np.int32))
    disc = cuda(np.array([1.0 / math.log(p + 2, 2) for p in range(k)], np.float64))
    hits = torch.empty((n_users, 1), dtype=torch.int32, device="cuda")
    dcg = torch.empty((n_users, 1), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hgr_rank_metrics(rd.data_ptr(), n_users, k, tp.data_ptr(), ti.data_ptr(), top_dev.data_ptr(), 1, disc.data_ptr(),
                                    hits.data_ptr(), dcg.data_ptr(), st))
    assert np.array_equal(hits[::997, 0].cpu().numpy(), hits_ref)
    idcg = [0.0]
    for j in range(20):
        idcg.append(idcg[-1] + 1.0 / math.log(j + 2, 2))
    out_h = torch.empty(1, dtype=torch.int64, device="cuda")
    out_r = torch.empty(1, dtype=torch.float64, device="cuda")
    out_n = torch.empty(1, dtype=torch.float64, device="cuda")
    for comp in (0, 1):
        _lib.check(lib.hgr_rank_metric_sums(hits.data_ptr(), dcg.data_ptr(), tp.data_ptr(), n_users, 1, top_dev.data_ptr(),
                                            cuda(np.array(idcg)).data_ptr(), len(idcg), comp, out_h.data_ptr(), out_r.data_ptr(),
                                            out_n.data_ptr(), st))
        h = hits[:, 0].cpu().numpy().tolist()
        terms = [a / b for a, b in zip(h, n_truth.tolist())]
        if comp == 0:
            want = 0.0
            for x in terms:
                want += x
        else:  # Neumaier, as CPython >= 3.12 sums floats
            want, c = 0.0, 0.0
            for x in terms:
                t = want + x
                c += (want - t) + x if abs(want) >= abs(x) else (x - t) + want
                want = t
            if c and math.isfinite(c):
                want += c
            if sys.version_info >= (3, 12):
                assert want == sum(terms)
        assert out_r.item() == want and out_h.item() == sum(h)
        nd = 0
        for d, m in zip(dcg[:, 0].cpu().numpy().tolist(), n_truth.tolist()):
            nd += d / idcg[min(m, 20)]
        assert out_n.item() == nd


def test_popular_items_with_low_ids_do_not_overflow_the_candidate_lists(E):
    """Ids are handed out in order of first appearance (data/ui_graph.py:43-68), so the popular items -- every user's best
    candidates after training -- share the first tiles of the catalogue: one (sub-range, column half) list must be able to take
    a user's whole candidate set.  40 000 items (so the catalogue is cut into many sub-ranges), the first 150 of them aligned
    with every user: exact results, and nobody falls back to the brute-force kernel."""
    rng = np.random.default_rng(77)
    n_users, n_items, k = 12_000, 40_000, 20
    common = rng.standard_normal(64).astype(np.float32)
    ue = (0.1 * rng.standard_normal((n_users, 64)) + 0.5 * common).astype(np.float32)
    ie = (0.1 * rng.standard_normal((n_items, 64))).astype(np.float32)
    ie[:150] += (0.5 * common * rng.uniform(0.5, 1.5, (150, 1))).astype(np.float32)
    tu = rng.integers(0, n_users, 100_000)
    ti = np.minimum((rng.pareto(1.2, 100_000) * 40).astype(np.int64), n_items - 1)  # training items concentrate on low ids too
    ptr, idx = train_csr(tu, ti, n_users)
    users = np.arange(0, n_users, 97)
    want_ids, want_sc = O.fullrank_topk(ue, ie, users, ptr, idx, k, mode="exact")
    all_users = np.arange(n_users, dtype=np.int32)
    ids, sc, stats = run(E, ue, ie, all_users, ptr, idx, k, "exact", "tensor")
    assert np.array_equal(ids[users], want_ids) and np.array_equal(sc[users].view(np.uint32), want_sc.view(np.uint32))
    assert stats[2] == 0, "%d users overflowed their candidate lists" % stats[2]
    assert (ids[:, :k] < 150).mean() > 0.9  # the case really is concentrated


def test_training_items_that_meet_the_threshold_never_reach_a_list(E):
    """The FILTER pass takes its group maxima over the raw scores and applies the training-item mask where a candidate is
    appended.  Users whose training items ARE their best-scoring items (embeddings out of a propagation), with training rows of
    1 - 200 items and of 300 - 4 000: always the oracle's lists, never a training item, and no list grows by the training row
    (no user of this graph may need the exact fallback)."""
    rng = np.random.default_rng(2024)
    n_users, n_items, d, k = 600, 9000, 64, 20
    degs = np.concatenate([rng.integers(1, 200, n_users - 6), [300, 500, 900, 1500, 2500, 4000]])
    tu = np.repeat(np.arange(n_users), degs)
    ti = np.concatenate([rng.choice(n_items, dg, replace=False) for dg in degs])
    ptr, idx = train_csr(tu, ti, n_users)
    ie = (rng.standard_normal((n_items, d)) * 0.1).astype(np.float32)
    ue = np.zeros((n_users, d), dtype=np.float32)
    for u in range(n_users):
        its = idx[ptr[u]:ptr[u + 1]]
        ue[u] = ie[its[:64]].mean(axis=0) * 3 + 0.02 * rng.standard_normal(d)
    users = np.arange(n_users, dtype=np.int32)
    want_ids, want_sc = O.fullrank_topk(ue, ie, users, ptr, idx, k, mode="exact")
    ids, sc, stats = run(E, ue, ie, users, ptr, idx, k, "exact", "tensor")
    assert np.array_equal(ids, want_ids) and np.array_equal(sc.view(np.uint32), want_sc.view(np.uint32))
    for u in (0, n_users - 6, n_users - 3, n_users - 1):
        assert not np.intersect1d(ids[u], idx[ptr[u]:ptr[u + 1]]).size
    assert stats[2] == 0, "%d users fell back to the brute-force kernel" % stats[2]
