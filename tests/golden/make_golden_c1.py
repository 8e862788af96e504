#!/usr/bin/env python
"""Golden vectors for BASELINE configs[0]: the hypergraph-diffusion (HGNN_HD3, --mode=local_only) encoder at the
lastfm SHAPE (1 891 users x 14 777 items, ~70 k training interactions, emb 64, 2 layers, fp32), produced by running
the UNMODIFIED reference on CPU.  The reference ships no dataset, so the graph is the synthetic power-law graph of that
shape; inputs that would be megabytes (the embedding table, the upstream gradient) come from an integer formula that any
host reproduces bit for bit (`formula_table`), and only what cannot be regenerated is stored: the encoder's parameters,
512 sampled output / gradient rows, column sums of the full outputs, the loss values and the reference's own top-20
lists + metric strings for 256 test users.

    python tests/golden/make_golden_c1.py        # needs /root/reference; writes tests/golden/c1_lastfm_shape.npz (~0.4 MB)
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import REF, install_shims, seed_all  # noqa: E402

N_USERS, N_ITEMS, N_TRAIN, SEED, D = 1891, 14777, 70_000, 2024, 64


def formula_table(rows: int, cols: int, salt: int, scale: float) -> np.ndarray:
    """Portable pseudo-random fp32 table: Knuth multiplicative hash of the element index, mapped to [-scale/2, scale/2)."""
    idx = np.arange(rows * cols, dtype=np.uint64) + np.uint64(salt) * np.uint64(1_000_003)
    h = (idx * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    h = ((h ^ (h >> np.uint64(15))) * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    return ((h.astype(np.float64) / 4294967296.0 - 0.5) * scale).astype(np.float32).reshape(rows, cols)


def graph_and_ids():
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions, reference_dense_ids

    g = powerlaw_interactions(N_USERS, N_ITEMS, N_TRAIN, seed=SEED)
    du, di, id2user, id2item = reference_dense_ids(g.train_u, g.train_i)
    return g, du, di, id2user, id2item


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not found at %s" % REF)
    install_shims()
    import torch

    torch.set_num_threads(1)
    work = tempfile.mkdtemp(prefix="hgr_golden_c1_")
    os.chdir(work)
    os.symlink(os.path.join(REF, "conf"), "conf")
    os.makedirs("log", exist_ok=True)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from base.graph_recommender import GraphRecommender
    from data.ui_graph import Interaction
    from util.evaluation import ranking_evaluation
    from util.loss_torch import bpr_loss, l2_reg_loss
    import model.graph.HGNN_HD3 as HD3

    g, du, di, id2user, id2item = graph_and_ids()
    train = [[int(u), int(i) + N_USERS, 1.0] for u, i in zip(g.train_u, g.train_i)]
    test = [[int(u), int(i) + N_USERS, 1.0] for u, i in zip(g.test_u, g.test_i)]
    data = Interaction(None, [list(t) for t in train], [list(t) for t in test])
    U, I = data.n_users, data.n_items
    N = U + I
    out = {"shape": np.array([N_USERS, N_ITEMS, N_TRAIN, SEED, D]), "n_users_items": np.array([U, I])}
    # the dense ids the reference assigned must be the ones synth.reference_dense_ids predicts (the test rebuilds them)
    assert np.array_equal(np.array([data.user[t[0]] for t in train]), du) and np.array_equal(np.array([data.item[t[1]] for t in train]), di)
    out["graph_checksum"] = np.array([int((du.astype(np.int64) * 31 + di.astype(np.int64)).sum()), int(du.size)])
    nadj = data.norm_adj.tocsr()
    out["norm_adj_nnz"] = np.array([nadj.nnz])
    out["norm_adj_value_sum"] = np.array([np.asarray(nadj.data, dtype=np.float64).sum()])
    out["norm_adj_row_sample"] = np.asarray(nadj[[0, 1, U - 1, U, N - 1]].todense(), dtype=np.float32)[:, :256]

    # ---- LocalAwareEncoder, 2 layers, eval mode (dropout off), non-trivial parameters
    seed_all(13)
    lae = HD3.LocalAwareEncoder(data, D, D, 2, 0.3, 0.2, torch.device("cpu"))
    lae.eval()
    for p in lae.parameters():
        p.data = p.data + 0.05 * torch.randn_like(p)
    for k, v in lae.state_dict().items():
        if v.numel() <= 64 * 64 + 64:  # the dense (U+I)^2 copies of the adjacency the reference keeps are not parameters
            out["lae_param/" + k] = v.numpy()
    E0 = torch.from_numpy(formula_table(N, D, 1, 0.2)).requires_grad_(True)
    G = torch.from_numpy(formula_table(N, D, 2, 2.0))
    lu, li = lae(E0, lae.sparse_norm_adj)
    full = torch.cat([lu, li], 0)
    (full * G).sum().backward()
    rows = np.arange(0, N, max(N // 512, 1))[:512]
    out["rows"] = rows
    out["lae_out_rows"] = full.detach().numpy()[rows]
    out["lae_out_colsum"] = full.detach().numpy().astype(np.float64).sum(0)
    out["lae_dE0_rows"] = E0.grad.numpy()[rows]
    out["lae_dE0_colsum"] = E0.grad.numpy().astype(np.float64).sum(0)
    for k, p in lae.named_parameters():
        if p.grad is not None and p.numel() <= 64 * 64 + 64:
            out["lae_grad/" + k] = p.grad.numpy()

    # ---- BPR + L2 on the encoder output with formula triples (LightGCN.py:52-55 / HGNN_HD3.py:339-343)
    B = 2048
    tri = (formula_table(B, 3, 3, 1.0) + 0.5)
    tu = np.minimum((tri[:, 0] * U).astype(np.int64), U - 1)
    tp = np.minimum((tri[:, 1] * I).astype(np.int64), I - 1)
    tn = np.minimum((tri[:, 2] * I).astype(np.int64), I - 1)
    ue, ie = lu.detach(), li.detach()
    rec = bpr_loss(ue[tu], ie[tp], ie[tn])
    reg = l2_reg_loss(0.01, ue[tu], ie[tp], ie[tn]) / B
    out["tri_u"], out["tri_p"], out["tri_n"] = tu, tp, tn
    out["loss_bpr"], out["loss_reg"] = rec.numpy(), reg.numpy()

    # ---- GraphRecommender.test on 256 test users + ranking_evaluation (top 10 / 20), the reference's own loop
    users = list(data.test_set.keys())[:256]
    sub = types.SimpleNamespace(test_set={u: data.test_set[u] for u in users}, user_rated=data.user_rated, item=data.item,
                                id2item=data.id2item, get_user_id=data.get_user_id)
    shell = types.SimpleNamespace(data=sub, max_N=20)
    shell.predict = lambda u: torch.matmul(ue[data.get_user_id(u)], ie.transpose(0, 1)).cpu().numpy()
    with contextlib.redirect_stdout(io.StringIO()):
        rec_list = GraphRecommender.test(shell)
    out["eval_users_raw"] = np.array(users)
    out["eval_rec_items_raw"] = np.array([[p[0] for p in rec_list[u]] for u in users])
    out["eval_rec_scores"] = np.array([[p[1] for p in rec_list[u]] for u in users], dtype=np.float32)
    out["eval_measures"] = np.array(ranking_evaluation(sub.test_set, rec_list, [10, 20]))

    path = os.path.join(HERE, "c1_lastfm_shape.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "(%d arrays, %.1f KB)" % (len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
