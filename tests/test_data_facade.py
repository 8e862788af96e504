"""Data layer façade (hypergraph_diffusion_for_recommendation_b200/data.py) against the reference's own ``FileIO`` /
``Interaction`` outputs (tests/golden/data_facade.{json,npz}, dumped by make_golden_data.py) and against the oracle's
python-loop restatement on random inputs.  Host-side views run on the CPU; the matrices are GPU tests (bit-exact)."""
import json
import os
import random

import numpy as np
import pytest

from hypergraph_diffusion_for_recommendation_b200 import data as D
from oracle import hgr_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "data_facade.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gold_arrays():
    return np.load(os.path.join(HERE, "golden", "data_facade.npz"))


def pairs(mapping):
    return [[k, v] for k, v in mapping.items()]


def nested(view):
    return [[k, pairs(view[k])] for k in view]


# ------------------------------------------------------------------------------------------ loader
def test_loader_matches_the_reference_on_tab_comma_and_mixed_files(gold, tmp_path):
    for name, text in gold["files"].items():
        p = tmp_path / name
        p.write_text(text)
        got = D.FileIO.load_data_set(str(p))
        assert isinstance(got, D.InteractionList)
        assert list(got) == gold["loaded"][name], name
        assert got == gold["loaded"][name] and len(got) == len(gold["loaded"][name])
        assert list(got) == O.load_data_set(str(p))  # the oracle's restatement agrees with the reference too
    empty = tmp_path / "empty.txt"
    empty.write_text("user\titem\n")
    assert len(D.FileIO.load_data_set(str(empty))) == 0
    bad = tmp_path / "bad.txt"
    bad.write_text("user\titem\n1\t2\n\n3\t4\n")  # a blank line: int('') raises in the reference
    with pytest.raises(ValueError):
        D.FileIO.load_data_set(str(bad))


def test_interaction_list_is_a_list_for_the_samplers():
    entries = [[5, 9, 1.0], [3, 9, 1.0], [5, 7, 1.0], [8, 1, 1.0]]
    il = D.InteractionList.from_entries(entries)
    assert len(il) == 4 and il[1] == [3, 9, 1.0] and list(il[1:3]) == entries[1:3]
    assert np.array_equal(np.array(il), np.array(entries))  # util/sampler.py:9
    a, b = list(entries), D.InteractionList.from_entries(entries)
    random.seed(3)
    random.shuffle(a)  # util/sampler.py:239 shuffles training_data in place
    random.seed(3)
    random.shuffle(b)
    assert list(b) == a
    assert isinstance(b[0][0], int) and isinstance(b[0][2], float)


# ------------------------------------------------------------------------------------------ Interaction, host views
@pytest.mark.parametrize("case", ["tab", "weighted"])
def test_interaction_views_match_the_reference(gold, gold_arrays, case):
    g = gold["cases"][case]
    d = D.Interaction(None, g["train"], g["test"])
    assert pairs(d.user) == g["user"] and pairs(d.item) == g["item"]
    assert pairs(d.id2user) == g["id2user"] and pairs(d.id2item) == g["id2item"]
    assert list(d.item.keys()) == [k for k, _ in g["item"]]  # util/sampler.py:251
    assert nested(d.training_set_u) == g["training_set_u"]
    assert nested(d.training_set_i) == g["training_set_i"]
    assert nested(d.test_set) == g["test_set"]
    assert [[u, d.user_history_dict[u]] for u in d.user_history_dict] == g["user_history_dict"]
    assert sorted(d.test_set_item) == g["test_set_item"]
    assert (d.n_users, d.n_items, d.n_cf_train, d.n_cf_test) == (g["n_users"], g["n_items"], g["n_cf_train"], g["n_cf_test"])
    assert list(d.training_size()) == g["training_size"] and list(d.test_size()) == g["test_size"]
    assert [[u, [list(x) for x in d.user_rated(u)]] for u, _ in g["user"]] == g["user_rated"]
    assert [[i, [list(x) for x in d.item_rated(i)]] for i, _ in g["item"]] == g["item_rated"]
    assert [[u, d.get_user_id(u)] for u, _ in g["get_user_id"]] == g["get_user_id"]
    assert [[i, d.get_item_id(i)] for i, _ in g["get_item_id"]] == g["get_item_id"]
    assert [[u, i, d.contain(u, i)] for u, i, _ in g["contain"]] == g["contain"]
    assert [[u, d.contain_user(u)] for u, _ in g["contain_user"]] == g["contain_user"]
    assert [[i, d.contain_item(i)] for i, _ in g["contain_item"]] == g["contain_item"]
    assert d.edge_index.tolist() == g["edge_index"] and d.edge_index_t.tolist() == g["edge_index_t"]
    assert np.array_equal(np.stack([d.row(k) for k in range(d.n_users)]), gold_arrays[case + "_row"])
    assert np.array_equal(np.stack([d.col(k) for k in range(d.n_items)]), gold_arrays[case + "_col"])
    assert np.array_equal(d.matrix(), gold_arrays[case + "_matrix"])
    # python types, not numpy scalars: the reference formats and hashes these
    u0 = next(iter(d.test_set))
    assert type(u0) is int and all(type(k) is int and type(v) is float for k, v in d.test_set[u0].items())
    assert d.training_set_u[424242] == {} and 424242 not in d.training_set_u  # defaultdict read of a missing user
    with pytest.raises(KeyError):
        d.user[424242]


def test_interaction_against_the_loop_restatement_on_random_lists():
    rng = np.random.default_rng(5)
    for trial in range(4):
        n = 4000
        # the last trial uses raw ids spread over 2^52: the packed-key sort does not fit and the np.unique path runs
        su, si = (1 << 52) // 300 if trial == 3 else 7, (1 << 52) // 600 if trial == 3 else 3
        tr = [[int(u), int(i), float(r)] for u, i, r in zip(rng.integers(0, 300, n) * su + 3, rng.integers(0, 500, n) * si + 11,
                                                            rng.choice([1.0, 1.0, 2.0, 0.5], n))]
        te = [[int(u), int(i), float(r)] for u, i, r in zip(rng.integers(0, 330, 900) * su + 3, rng.integers(0, 600, 900) * si + 11,
                                                            rng.choice([1.0, 3.0], 900))]
        ref = O.generate_set(tr, te)
        d = D.Interaction(None, tr, te)
        assert dict(d.user) == ref["user"] and list(d.user) == list(ref["user"])
        assert dict(d.item) == ref["item"] and list(d.item) == list(ref["item"])
        assert dict(d.id2user) == ref["id2user"] and dict(d.id2item) == ref["id2item"]
        for view, want in ((d.training_set_u, ref["training_set_u"]), (d.training_set_i, ref["training_set_i"]), (d.test_set, ref["test_set"])):
            assert list(view) == list(want)
            assert all(pairs(view[k]) == pairs(want[k]) for k in want)
        assert list(d.user_history_dict) == list(ref["user_history_dict"])
        assert all(d.user_history_dict[k] == v for k, v in ref["user_history_dict"].items())
        assert d.test_set_item == ref["test_set_item"]
        du, di = d.dense_training_pairs()
        assert du.tolist() == [ref["user"][e[0]] for e in tr] and di.tolist() == [ref["item"][e[1]] for e in tr]
        tu, raw, ptr, truth, id2item = d.eval_arrays()
        assert raw == list(ref["test_set"]) and tu.tolist() == [ref["user"][u] for u in raw]
        assert truth.tolist() == [ref["item"].get(i, -1) for u in raw for i in ref["test_set"][u]]
        assert np.array_equal(np.diff(ptr), [len(ref["test_set"][u]) for u in raw])
        assert id2item.tolist() == [ref["id2item"][k] for k in range(len(ref["item"]))]


@pytest.mark.parametrize("case", ["tab", "weighted"])
def test_oracle_matrices_match_the_reference(gold, gold_arrays, case):
    """The oracle's builders against the reference's five matrices, including the shape-picked branch of
    normalize_graph_mat (tab: 5 users x 5 items -> the square branch on the interaction matrix)."""
    g = gold["cases"][case]
    d = D.Interaction(None, g["train"], g["test"])
    u, i = d.dense_training_pairs()
    nu, ni = d.n_users, d.n_items

    def same(got, name):
        for a, key in zip(got, ("indptr", "indices", "values")):
            want = gold_arrays["%s_%s_%s" % (case, name, key)]
            assert np.array_equal(np.asarray(a, dtype=want.dtype).view(np.uint32 if key == "values" else want.dtype),
                                  want.view(np.uint32 if key == "values" else want.dtype)), (name, key)

    same(O.bipartite_adjacency(u, i, nu, ni), "ui_adj")
    same(O.build_norm_adj(u, i, nu, ni), "norm_adj")
    r = O.interaction_matrix(u, i, nu, ni)
    rt = O.interaction_matrix(i, u, ni, nu)
    same(r, "interaction_mat")
    same(rt, "inv_interaction_mat")
    same(O.normalize_graph_mat(*r, ni), "norm_interaction_mat")
    same(O.normalize_graph_mat(*rt, nu), "norm_inv_interaction_mat")


def test_matrices_fail_loudly_without_a_gpu(gold):
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a host without CUDA")
    d = D.Interaction(None, gold["cases"]["tab"]["train"], gold["cases"]["tab"]["test"])
    from hypergraph_diffusion_for_recommendation_b200._lib import HgrError

    with pytest.raises(HgrError):
        d.norm_adj


# ------------------------------------------------------------------------------------------ Interaction, device matrices
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tab", "weighted"])
def test_device_matrices_are_bit_identical_to_the_reference(gold, gold_arrays, case):
    g = gold["cases"][case]
    d = D.Interaction(None, g["train"], g["test"])
    shapes = {"ui_adj": (d.n_users + d.n_items,) * 2, "norm_adj": (d.n_users + d.n_items,) * 2,
              "interaction_mat": (d.n_users, d.n_items), "inv_interaction_mat": (d.n_items, d.n_users),
              "norm_interaction_mat": (d.n_users, d.n_items), "norm_inv_interaction_mat": (d.n_items, d.n_users)}
    for name, shape in shapes.items():
        m = getattr(d, name)
        assert m.shape == shape and getattr(d, name) is m  # built once
        ip, ix, dv = m.to_host()
        assert np.array_equal(ip, gold_arrays["%s_%s_indptr" % (case, name)]), name
        assert np.array_equal(ix, gold_arrays["%s_%s_indices" % (case, name)]), name
        assert np.array_equal(dv.view(np.uint32), gold_arrays["%s_%s_values" % (case, name)].view(np.uint32)), name
    # Graph.normalize_graph_mat on a handle: same bits as the matrices the constructor normalises
    for raw, normed in (("ui_adj", "norm_adj"), ("interaction_mat", "norm_interaction_mat"), ("inv_interaction_mat", "norm_inv_interaction_mat")):
        got = d.normalize_graph_mat(getattr(d, raw)).to_host()
        assert np.array_equal(got[2].view(np.uint32), gold_arrays["%s_%s_values" % (case, normed)].view(np.uint32))


@pytest.mark.gpu
def test_convert_to_laplacian_mat_of_a_perturbed_graph(gold, gold_arrays):
    import torch

    g = gold["cases"]["tab"]
    d = D.Interaction(None, g["train"], g["test"])
    m = d.interaction_mat
    vals = m.values.clone()
    ip = m.indptr.cpu().numpy()
    ix = m.indices.cpu().numpy()
    for r, c in gold_arrays["tab_laplacian_dropped"]:
        p = ip[r] + int(np.nonzero(ix[ip[r]:ip[r + 1]] == c)[0][0])
        vals[p] = 0
    lap = d.convert_to_laplacian_mat(m.with_values(vals)).to_host()
    assert np.array_equal(lap[0], gold_arrays["tab_laplacian_indptr"])
    assert np.array_equal(lap[1], gold_arrays["tab_laplacian_indices"])
    assert np.array_equal(lap[2].view(np.uint32), gold_arrays["tab_laplacian_values"].view(np.uint32))
    assert torch.is_tensor(d.edge_index)


@pytest.mark.gpu
def test_eval_and_sampler_take_the_facade_without_python_loops(gold):
    import torch

    from hypergraph_diffusion_for_recommendation_b200 import evaluation, sampler

    g = gold["cases"]["tab"]
    d = D.Interaction(None, g["train"], g["test"])
    ref = O.generate_set(g["train"], g["test"])
    ev = evaluation.EvalData(d)
    assert ev.raw_users == list(ref["test_set"])
    assert ev.test_users.tolist() == [ref["user"][u] for u in ev.raw_users]
    tr = sorted({(ref["user"][u], ref["item"][i]) for u, i, _ in g["train"]})
    ptr = ev.train_indptr.cpu().numpy()
    assert [(r, int(c)) for r in range(d.n_users) for c in ev.train_indices.cpu().numpy()[ptr[r]:ptr[r + 1]]] == tr
    assert ev.truth_items.tolist() == [ref["item"].get(i, -1) for u in ev.raw_users for i in ref["test_set"][u]]
    assert ev.id2item.tolist() == [ref["id2item"][k] for k in range(d.n_items)]
    s = sampler.PairwiseSampler.from_interaction(d)
    assert s.edge_u.tolist() == [ref["user"][e[0]] for e in g["train"]] and s.edge_i.tolist() == [ref["item"][e[1]] for e in g["train"]]
    for u, p, n in sampler.next_batch_pairwise(d, 5):
        assert u.dtype == torch.int64 and u.numel() == p.numel() == n.numel()
        for uu, pp, nn in zip(u.tolist(), p.tolist(), n.tolist()):
            assert (uu, pp) in tr and (uu, nn) not in tr


def test_empty_and_single_interaction_lists():
    d = D.Interaction(None, [], [])
    assert (d.n_users, d.n_items) == (0, 0) and d.training_size() == (0, 0, 0) and d.test_size() == (0, 0, 0)
    assert list(d.test_set) == [] and d.user_rated(5) == ([], []) and d.contain(1, 2) is False and d.get_user_id(1) is None
    d = D.Interaction(None, [[1, 2, 1.0]], [[1, 3, 1.0], [9, 9, 1.0]])  # the test entry of the unknown user 9 is skipped
    assert d.training_size() == (1, 1, 1) and d.test_size() == (1, 1, 2)
    assert dict(d.test_set[1]) == {3: 1.0} and d.user_rated(1) == ([2], [1.0]) and d.test_set_item == {3}
    users, raw, ptr, truth, id2item = d.eval_arrays()
    assert users.tolist() == [0] and raw == [1] and ptr.tolist() == [0, 1] and truth.tolist() == [-1] and id2item.tolist() == [2]
