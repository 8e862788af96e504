"""BASELINE configs[0]: the hypergraph-diffusion (HGNN_HD3 local) encoder at the lastfm shape, against golden vectors the
UNMODIFIED reference produced on CPU (tests/golden/make_golden_c1.py).  The CPU test pins the oracle, the GPU test the
CUDA path through the C ABI: adjacency, encoder forward / backward, BPR + L2, full-rank top-20 lists and metric strings."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_c1 import D, N_ITEMS, N_USERS, formula_table, graph_and_ids  # noqa: E402

from oracle import hgr_oracle as O  # noqa: E402

RTOL = 1e-5


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def c1():
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_lastfm_shape.npz"))
    g, du, di, id2user, id2item = graph_and_ids()
    assert [int((du.astype(np.int64) * 31 + di.astype(np.int64)).sum()), int(du.size)] == gold["graph_checksum"].tolist(), \
        "the synthetic generator no longer reproduces the graph the golden vectors were made from"
    U, I = (int(x) for x in gold["n_users_items"])
    params = {k[len("lae_param/"):]: gold[k] for k in gold.files if k.startswith("lae_param/")}
    return dict(gold=gold, g=g, du=du, di=di, id2user=id2user, id2item=id2item, U=U, I=I, N=U + I, params=params,
                E0=formula_table(U + I, D, 1, 0.2), G=formula_table(U + I, D, 2, 2.0))


def check_adjacency(c1, indptr, indices, values):
    gold, U, N = c1["gold"], c1["U"], c1["N"]
    assert indices.size == int(gold["norm_adj_nnz"][0])
    assert abs(values.astype(np.float64).sum() - float(gold["norm_adj_value_sum"][0])) < 1e-6 * float(gold["norm_adj_value_sum"][0])
    for k, r in enumerate([0, 1, U - 1, U, N - 1]):
        dense = np.zeros(N, dtype=np.float32)
        dense[indices[indptr[r]:indptr[r + 1]]] = values[indptr[r]:indptr[r + 1]]
        assert np.array_equal(dense[:256].view(np.uint32), gold["norm_adj_row_sample"][k].view(np.uint32)), r


def assert_lists_equal_up_to_rounding_ties(ids, sc, ref_ids, ref_sc, tol=2e-6):
    """The reference scores with a BLAS GEMV (summation order unspecified) and this path with the canonical FMA chain, so
    two items whose scores agree to rounding may swap places.  Everything else must be identical: scores position by
    position, and the ids inside every group of positions whose reference scores are separated by more than `tol`
    (relative) from their neighbours; only a group that touches the cut-off K may exchange an id with the outside."""
    k = ref_ids.shape[1]
    assert np.abs(sc - ref_sc).max() <= tol * np.abs(ref_sc).max()
    exact = 0
    for r in range(ref_ids.shape[0]):
        if np.array_equal(ids[r], ref_ids[r]):
            exact += 1
            continue
        gaps = np.abs(np.diff(ref_sc[r])) > tol * np.abs(ref_sc[r][:-1])
        cuts = [0] + (np.nonzero(gaps)[0] + 1).tolist() + [k]
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b < k:
                assert sorted(ids[r][a:b]) == sorted(ref_ids[r][a:b]), (r, a, b, ids[r], ref_ids[r])
    return exact / ref_ids.shape[0]


def truth_and_users(c1):
    """Test users (reference order = first appearance in the test file, users seen in training only) -> dense ids."""
    gold, g = c1["gold"], c1["g"]
    user = {int(r): k for k, r in enumerate(c1["id2user"])}
    item = {int(r): k for k, r in enumerate(c1["id2item"])}
    users_raw = [int(x) for x in gold["eval_users_raw"]]
    truth = {u: [] for u in users_raw}
    for uu, ii in zip(g.test_u.tolist(), g.test_i.tolist()):
        if uu in truth:
            truth[uu].append(item.get(ii, -1))
    ptr = np.zeros(len(users_raw) + 1, np.int64)
    np.cumsum([len(truth[u]) for u in users_raw], out=ptr[1:])
    items = np.array([x for u in users_raw for x in truth[u]], dtype=np.int64)
    return np.array([user[u] for u in users_raw]), ptr, items


def test_oracle_matches_the_reference_at_the_lastfm_shape(c1):
    gold, U, I, N = c1["gold"], c1["U"], c1["I"], c1["N"]
    assert U == N_USERS and I <= N_ITEMS  # items that only occur in the held-out split get no dense id
    csr = O.build_norm_adj(c1["du"], c1["di"], U, I)
    check_adjacency(c1, *csr)
    lu, li = O.local_aware_encoder(csr, c1["E0"], c1["params"], 2, U)
    full = np.concatenate([lu, li])
    assert rel(full[gold["rows"]], gold["lae_out_rows"]) < RTOL
    assert rel(full.astype(np.float64).sum(0), gold["lae_out_colsum"]) < 1e-4
    rec, reg, _, _ = O.bpr_l2_from_tables(lu, li, gold["tri_u"], gold["tri_p"], gold["tri_n"], 0.01, 2048)
    assert rel(rec, gold["loss_bpr"]) < RTOL and rel(reg, gold["loss_reg"]) < RTOL
    test_users, ptr, items = truth_and_users(c1)
    tip, tix, _ = O.interaction_matrix(c1["du"], c1["di"], U, I)
    ids, sc = O.fullrank_topk(lu, li, test_users, tip, tix, 20, mode="refquirk")
    raw = c1["id2item"][ids] + N_USERS  # the files carry item ids offset by the number of users
    assert assert_lists_equal_up_to_rounding_ties(raw, sc, gold["eval_rec_items_raw"], gold["eval_rec_scores"]) > 0.95
    from hypergraph_diffusion_for_recommendation_b200.evaluation import ranking_evaluation_ids

    assert ranking_evaluation_ids(ptr, items, ids, [10, 20]) == [str(s) for s in gold["eval_measures"]]


@pytest.mark.gpu
def test_cuda_path_matches_the_reference_at_the_lastfm_shape(c1):
    import types

    import torch

    from hypergraph_diffusion_for_recommendation_b200 import _lib, encoders, evaluation, graph, loss_torch

    _lib.lib()
    gold, U, I, N = c1["gold"], c1["U"], c1["I"], c1["N"]
    dev = torch.device("cuda", 0)
    adj = graph.build_norm_adj(c1["du"], c1["di"], U, I, device=dev)
    check_adjacency(c1, *adj.to_host())
    data = types.SimpleNamespace(n_users=U, n_items=I, norm_adj=None, norm_adj_device=adj)
    lae = encoders.LocalAwareEncoder(data, D, D, 2, 0.3, 0.2, dev).to(dev)
    missing, unexpected = lae.load_state_dict({k: torch.from_numpy(v) for k, v in c1["params"].items()}, strict=False)
    assert not unexpected and all("hgnn_layers" in k or "hgcn_layer" in k for k in missing), (missing, unexpected)
    lae.eval()
    e0 = torch.from_numpy(c1["E0"]).to(dev).requires_grad_(True)
    lu, li = lae(e0, adj)
    full = torch.cat([lu, li], 0)
    rows = torch.from_numpy(gold["rows"]).to(dev)
    assert rel(full[rows].detach().cpu().numpy(), gold["lae_out_rows"]) < RTOL
    assert rel(full.detach().double().sum(0).cpu().numpy(), gold["lae_out_colsum"]) < 1e-4
    (full * torch.from_numpy(c1["G"]).to(dev)).sum().backward()
    assert rel(e0.grad[rows].cpu().numpy(), gold["lae_dE0_rows"]) < 5e-5
    assert rel(e0.grad.double().sum(0).cpu().numpy(), gold["lae_dE0_colsum"]) < 1e-4
    grads = {k: p.grad for k, p in lae.named_parameters() if p.grad is not None}
    for k in gold.files:
        if k.startswith("lae_grad/"):
            assert rel(grads[k[len("lae_grad/"):]].cpu().numpy(), gold[k]) < 2e-4, k
    rec, reg = loss_torch.bpr_l2_from_tables(lu.detach(), li.detach(), torch.from_numpy(gold["tri_u"]), torch.from_numpy(gold["tri_p"]),
                                             torch.from_numpy(gold["tri_n"]), 0.01, 2048)
    assert rel(rec.cpu().numpy(), gold["loss_bpr"]) < RTOL and rel(reg.cpu().numpy(), gold["loss_reg"]) < RTOL
    # full-rank evaluation of the reference's 256 test users, the reference's find_k_largest semantics
    test_users, ptr, items = truth_and_users(c1)
    train = graph.build_interaction_csr(c1["du"], c1["di"], U, I, device=dev)
    for engine in ("tensor", "simt"):
        ids, sc = evaluation.fullrank_topk(lu.detach(), li.detach(), torch.from_numpy(test_users).to(dev), train.indptr, train.indices, 20,
                                           mode="refquirk", engine=engine)
        ids_h = ids.cpu().numpy()
        assert assert_lists_equal_up_to_rounding_ties(c1["id2item"][ids_h] + N_USERS, sc.cpu().numpy(), gold["eval_rec_items_raw"],
                                                      gold["eval_rec_scores"]) > 0.95, engine
        assert evaluation.ranking_evaluation_ids(ptr, items, ids_h, [10, 20]) == [str(s) for s in gold["eval_measures"]]
        assert evaluation.ranking_evaluation_device(ptr, items, ids, [10, 20]) == [str(s) for s in gold["eval_measures"]]
