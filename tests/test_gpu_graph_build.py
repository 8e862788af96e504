"""GPU parity of the device graph builder (csrc/graph_build.cu) -- bit-exact against the oracle's
restatement of the reference's scipy path, which tests/test_oracle_golden.py pins to the reference."""
import numpy as np
import pytest
import torch

from oracle import hgr_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from hypergraph_diffusion_for_recommendation_b200 import graph

    return graph


def same_csr(dev_csr, want, values_bit_exact=True):
    ip, ix, dv = dev_csr.to_host()
    assert np.array_equal(ip, want[0]), "indptr differs"
    assert np.array_equal(ix.astype(np.int64), want[1]), "indices differ"
    if values_bit_exact:
        assert np.array_equal(dv.view(np.uint32), np.asarray(want[2], dtype=np.float32).view(np.uint32)), "values differ"


def test_hand_graph_with_duplicate_and_structure_vs_golden(G, golden):
    tr = golden["hand_train"]
    user = {r: k for k, r in enumerate(golden["hand_id2user"])}
    item = {r: k for k, r in enumerate(golden["hand_id2item"])}
    u = np.array([user[int(t[0])] for t in tr])
    i = np.array([item[int(t[1])] for t in tr])
    raw = G.build_norm_adj(u, i, 3, 4, normalize=False)
    same_csr(raw, (golden["hand_ui_indptr"], golden["hand_ui_indices"], golden["hand_ui_data"]))  # duplicate summed to 2.0
    adj = G.build_norm_adj(u, i, 3, 4)
    same_csr(adj, O.build_norm_adj(u, i, 3, 4))
    same_csr(adj, (golden["hand_norm_indptr"], golden["hand_norm_indices"], None), values_bit_exact=False)
    assert np.abs(adj.to_host()[2] - golden["hand_norm_data"]).max() < 1e-7
    assert list(adj.degree.cpu().numpy()) == [3, 2, 3, 2, 3, 1, 2]


def test_powerlaw_graph_bit_exact(G, golden, pl_graph):
    adj = G.build_norm_adj(pl_graph["u"], pl_graph["i"], pl_graph["n_users"], pl_graph["n_items"])
    same_csr(adj, pl_graph["csr"])
    same_csr(adj, (golden["pl_norm_indptr"], golden["pl_norm_indices"], None), values_bit_exact=False)
    assert adj.symmetric and adj.t() is adj


@pytest.mark.parametrize("n_edges", [1, 4095, 4096, 4097, 70_001, 300_000])
def test_random_graphs_with_duplicates(G, n_edges):
    rng = np.random.default_rng(n_edges)
    n_users, n_items = 700, 1300
    u = rng.integers(0, n_users, n_edges)
    i = (rng.zipf(1.3, n_edges) - 1) % n_items  # heavy duplicates on the popular items
    adj = G.build_norm_adj(u, i, n_users, n_items)
    want = O.build_norm_adj(u, i, n_users, n_items)
    same_csr(adj, want)
    r = G.build_interaction_csr(u, i, n_users, n_items)
    same_csr(r, O.interaction_matrix(u, i, n_users, n_items))
    rt = G.build_interaction_csr(u, i, n_users, n_items, transpose=True)
    same_csr(rt, O.interaction_matrix(i, u, n_items, n_users))
    rn = G.build_interaction_csr(u, i, n_users, n_items, row_normalize=True)
    same_csr(rn, O.normalize_graph_mat(*O.interaction_matrix(u, i, n_users, n_items), n_items))


def test_empty_and_isolated_nodes(G):
    adj = G.build_norm_adj(np.zeros(0, np.int64), np.zeros(0, np.int64), 5, 6)
    assert adj._nnz() == 0 and not adj.indptr.any()
    # users 1, 3 and items 0, 2 never interact: empty rows, degree 0 -> scale 0 (inf -> 0 in the reference)
    u, i = np.array([0, 2, 2, 4]), np.array([1, 1, 3, 3])
    adj = G.build_norm_adj(u, i, 5, 4)
    same_csr(adj, O.build_norm_adj(u, i, 5, 4))
    assert list(adj.degree.cpu().numpy()) == [1, 0, 2, 0, 1, 0, 2, 0, 2]


def test_properties_at_c3_scale(G):
    """30 k x 41 k x 1 M interactions (BASELINE config C3): canonical form + symmetry + degree sums,
    and value parity with the oracle on a row sample."""
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions_device

    U, I, E = 30_000, 41_000, 1_000_000
    u, i = powerlaw_interactions_device(U, I, E, torch.device("cuda"), seed=5)
    adj = G.build_norm_adj(u, i, U, I)
    ip, ix, dv = adj.to_host()
    assert adj._nnz() == 2 * E and ip[-1] == 2 * E
    rows = np.repeat(np.arange(U + I), np.diff(ip))
    key = rows.astype(np.int64) * (U + I) + ix
    assert np.all(np.diff(key) > 0)  # sorted, no duplicates
    tkey = np.sort(ix.astype(np.int64) * (U + I) + rows)
    assert np.array_equal(key, tkey)  # structurally symmetric
    assert np.array_equal(adj.degree.cpu().numpy(), np.diff(ip))
    d = G.host_pow_lut(int(np.diff(ip).max()) + 1, -0.5)[np.diff(ip)]
    assert np.array_equal(dv.view(np.uint32), (d[rows] * d[ix]).astype(np.float32).view(np.uint32))
    want = O.build_norm_adj(u.cpu().numpy(), i.cpu().numpy(), U, I)
    assert np.array_equal(ip, want[0]) and np.array_equal(ix, want[1])
