"""The reference's LightGCN training loop (model/graph/LightGCN.py:36-66) replayed through the drop-in path end to end:
``data.Interaction`` (façade) -> ``encoders.LGCN_Encoder`` -> ``loss_torch.bpr_l2_from_tables`` -> Adam -> ``evaluation.test`` ->
``evaluation.ranking_evaluation``, against the trajectory the reference's own classes produced on the CPU
(tests/golden/lightgcn_loop.npz, make_golden_loop.py): same sampled triples; per-batch losses within 1e-5 relative (north_star), per-epoch
tables within 1e-4 (Adam amplifies rounding on rarely sampled rows), metric values of every epoch, and identical metric strings on the
reference's own tables."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


from conftest import rowwise_rel_err as rel  # noqa: E402  row-wise: each embedding row against its own magnitude


def test_lightgcn_trajectory_matches_the_reference():
    from hypergraph_diffusion_for_recommendation_b200 import data as D
    from hypergraph_diffusion_for_recommendation_b200 import encoders, evaluation, loss_torch

    g = np.load(os.path.join(HERE, "golden", "lightgcn_loop.npz"))
    emb, layers, batch, epochs, lr, wdecay, reg = g["conf"]
    emb, layers, batch, epochs = int(emb), int(layers), int(batch), int(epochs)
    data = D.Interaction(None, g["train"].tolist(), g["test"].tolist())
    model = encoders.LGCN_Encoder(data, emb, layers).cuda()
    with torch.no_grad():
        model.embedding_dict["user_emb"].copy_(torch.from_numpy(g["init_user_emb"]))
        model.embedding_dict["item_emb"].copy_(torch.from_numpy(g["init_item_emb"]))
    optimizer = torch.optim.Adam(model.parameters(), lr=float(lr), weight_decay=float(wdecay))
    rec = types.SimpleNamespace(data=data, max_N=20)
    triples = torch.from_numpy(g["triples"]).cuda()
    off, b = 0, 0
    for epoch in range(epochs):
        while b < g["batch_sizes"].size and g["batch_epoch"][b] == epoch:
            n = int(g["batch_sizes"][b])
            u, p, ng = (triples[k, off:off + n] for k in range(3))
            ue, ie = model()
            rec_loss, reg_loss = loss_torch.bpr_l2_from_tables(ue, ie, u, p, ng, float(reg), batch)
            optimizer.zero_grad()
            (rec_loss + reg_loss).backward()
            optimizer.step()
            want = g["losses"][b]
            assert abs(rec_loss.item() - want[0]) <= 1e-5 * abs(want[0]), (b, rec_loss.item(), want[0])
            assert abs(reg_loss.item() - want[1]) <= 1e-5 * abs(want[1]), (b, reg_loss.item(), want[1])
            off += n
            b += 1
        with torch.no_grad():
            ue, ie = model()
        # Adam divides by sqrt(v) + 1e-8: rounding differences of the first steps are amplified for rows that are hardly ever
        # sampled, so the parameters get 1e-4, the propagated tables (what the recommender uses) 1e-5 relative to the table's scale
        assert rel(model.embedding_dict["user_emb"], g["epoch%d_user_param" % epoch]) < 1e-4
        assert rel(model.embedding_dict["item_emb"], g["epoch%d_item_param" % epoch]) < 1e-4
        assert rel(ue, g["epoch%d_user_emb" % epoch]) < 1e-4 and rel(ie, g["epoch%d_item_emb" % epoch]) < 1e-4
        # the FORWARD pass on the reference's own parameters of this epoch: 1e-5 row-wise (north_star), no optimiser in between
        probe = encoders.LGCN_Encoder(data, emb, layers).cuda()
        with torch.no_grad():
            probe.embedding_dict["user_emb"].copy_(torch.from_numpy(g["epoch%d_user_param" % epoch]))
            probe.embedding_dict["item_emb"].copy_(torch.from_numpy(g["epoch%d_item_param" % epoch]))
            pu, pi = probe()
        assert rel(pu, g["epoch%d_user_emb" % epoch]) < 1e-5 and rel(pi, g["epoch%d_item_emb" % epoch]) < 1e-5
        rec_list = evaluation.test(rec, ue, ie)  # refquirk mode: the reference's find_k_largest semantics
        got = evaluation.ranking_evaluation(data.test_set, rec_list, [10, 20])
        want = [str(s) for s in g["epoch%d_measures" % epoch]]
        # scores agree to ~1e-6, so a near-tie at a list boundary may swap two items for a user: the metric VALUES must
        # agree to 2e-3; on the reference's own tables the strings are identical (tests/test_gpu_eval.py)
        assert len(got) == len(want)
        for a, w in zip(got, want):
            if ":" in w:
                assert a.split(":")[0] == w.split(":")[0]
                assert abs(float(a.split(":")[1]) - float(w.split(":")[1])) <= 2e-3, (a, w)
            else:
                assert a == w
        # the same evaluation on the reference's tables of this epoch: identical strings
        ref_list = evaluation.test(rec, torch.from_numpy(g["epoch%d_user_emb" % epoch]).cuda(), torch.from_numpy(g["epoch%d_item_emb" % epoch]).cuda())
        assert evaluation.ranking_evaluation(data.test_set, ref_list, [10, 20]) == want
    assert b == g["batch_sizes"].size


def test_files_to_metrics_through_the_public_api(tmp_path):
    """train.txt / test.txt -> ``FileIO.load_data_set`` -> ``Interaction`` -> ``LGCN_Encoder`` -> device sampler
    (``next_batch_pairwise``) -> fused loss -> Adam -> ``evaluation.test`` -> ``ranking_evaluation``: the reference's
    SELFRec.execute path (SELFRec.py:15-33, model/graph/LightGCN.py:36-102) end to end on this package's modules.  Two epochs
    on a synthetic power-law split must beat the popularity-free starting point by a wide margin."""
    from hypergraph_diffusion_for_recommendation_b200 import data as D
    from hypergraph_diffusion_for_recommendation_b200 import encoders, evaluation, loss_torch, sampler
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions, write_reference_files

    g = powerlaw_interactions(500, 800, 16000, seed=9)
    root = write_reference_files(g, str(tmp_path), dataset="lastfm")
    train = D.FileIO.load_data_set(os.path.join(root, "train.txt"))
    test = D.FileIO.load_data_set(os.path.join(root, "test.txt"))
    assert len(train) == g.train_u.size and len(test) == g.test_u.size
    data = D.Interaction(None, train, test)
    assert data.training_size() == (data.n_users, data.n_items, len(train))
    torch.manual_seed(0)
    model = encoders.LGCN_Encoder(data, 64, 3).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    rec = types.SimpleNamespace(data=data, max_N=20)

    def recall20():
        with torch.no_grad():
            ue, ie = model()
        lines = evaluation.ranking_evaluation(data.test_set, evaluation.test(rec, ue, ie), [20])
        return float([s for s in lines if s.startswith("Recall")][0].split(":")[1])

    before = recall20()
    for epoch in range(2):
        n_batches = 0
        for u, p, n in sampler.next_batch_pairwise(data, 2048):
            ue, ie = model()
            rec_loss, reg_loss = loss_torch.bpr_l2_from_tables(ue, ie, u, p, n, 1e-4, 2048)
            opt.zero_grad()
            (rec_loss + reg_loss).backward()
            opt.step()
            n_batches += 1
        assert n_batches == -(-len(train) // 2048)  # the short last batch included, as in the reference
    after = recall20()
    assert torch.isfinite(rec_loss) and after > 2 * before and after > 0.1, (before, after)
