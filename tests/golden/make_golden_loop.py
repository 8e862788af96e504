#!/usr/bin/env python
"""Golden TRAJECTORY of the reference's LightGCN training loop (model/graph/LightGCN.py:36-66), produced by driving the
reference's own components exactly as ``LightGCN.train`` does: ``Interaction`` -> ``LGCN_Encoder`` -> per batch
``next_batch_pairwise`` / ``bpr_loss`` / ``l2_reg_loss`` / Adam -> per epoch ``GraphRecommender.test`` + ``ranking_evaluation``.

    python tests/golden/make_golden_loop.py      # needs /root/reference; rewrites tests/golden/lightgcn_loop.npz

Stored: the interaction lists, the initial tables, every batch's (user, pos, neg) dense ids as the reference's sampler drew
them, every batch's rec / reg loss, the tables after each epoch, and the metric strings of each epoch.  The GPU test
(tests/test_gpu_training_loop.py) replays the same triples through the façade + libhgr path and must land on the same numbers.
"""
import os
import random
import sys
import tempfile
import types
import warnings

import numpy as np

REF = os.environ.get("HGR_REFERENCE", "/root/reference/HD_SELFRec")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

N_USERS, N_ITEMS, N_TRAIN = 300, 400, 6000
EMB, LAYERS, BATCH, EPOCHS = 64, 3, 1024, 2
LR, WDECAY, REG = 0.005, 5e-6, 0.1
TOPN = [10, 20]


def main():
    import torch

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.set_num_threads(1)
    work = tempfile.mkdtemp(prefix="hgr_golden_loop_")
    os.chdir(work)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from base.graph_recommender import GraphRecommender
    from data.ui_graph import Interaction
    from model.graph.LightGCN import LGCN_Encoder
    from util.evaluation import ranking_evaluation
    from util.loss_torch import bpr_loss, l2_reg_loss
    from util.sampler import next_batch_pairwise

    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(N_USERS, N_ITEMS, N_TRAIN, seed=21)
    # raw ids that are not the dense ids: users 1000 + 3 u, items 50000 + 7 i, file order shuffled
    rng = np.random.default_rng(4)
    order = rng.permutation(g.train_u.size)
    train = [[int(1000 + 3 * g.train_u[k]), int(50000 + 7 * g.train_i[k]), 1.0] for k in order]
    test = [[int(1000 + 3 * u), int(50000 + 7 * i), 1.0] for u, i in zip(g.test_u, g.test_i)]
    data = Interaction(None, [list(e) for e in train], [list(e) for e in test])

    random.seed(20)
    np.random.seed(20)
    torch.manual_seed(20)
    model = LGCN_Encoder(data, EMB, LAYERS)
    out = {"train": np.array(train), "test": np.array(test), "conf": np.array([EMB, LAYERS, BATCH, EPOCHS, LR, WDECAY, REG]),
           "init_user_emb": model.embedding_dict["user_emb"].detach().numpy().copy(),
           "init_item_emb": model.embedding_dict["item_emb"].detach().numpy().copy()}
    optimizer = torch.optim.Adam(model.parameters(), lr=LR, weight_decay=WDECAY)
    shell = types.SimpleNamespace(data=data, max_N=max(TOPN))
    triples, losses, batch_epoch = [], [], []
    for epoch in range(EPOCHS):
        for n, batch in enumerate(next_batch_pairwise(data, BATCH)):  # model/graph/LightGCN.py:49-66
            user_idx, pos_idx, neg_idx = batch
            rec_user_emb, rec_item_emb = model()
            user_emb, pos_item_emb, neg_item_emb = rec_user_emb[user_idx], rec_item_emb[pos_idx], rec_item_emb[neg_idx]
            rec_loss = bpr_loss(user_emb, pos_item_emb, neg_item_emb)
            reg_loss = l2_reg_loss(REG, user_emb, pos_item_emb, neg_item_emb) / BATCH
            batch_loss = rec_loss + reg_loss
            optimizer.zero_grad()
            batch_loss.backward()
            optimizer.step()
            triples.append(np.stack([np.asarray(user_idx), np.asarray(pos_idx), np.asarray(neg_idx)]))
            losses.append([rec_loss.item(), reg_loss.item()])
            batch_epoch.append(epoch)
        with torch.no_grad():
            ue, ie = model()
        out["epoch%d_user_emb" % epoch], out["epoch%d_item_emb" % epoch] = ue.numpy().copy(), ie.numpy().copy()
        out["epoch%d_user_param" % epoch] = model.embedding_dict["user_emb"].detach().numpy().copy()
        out["epoch%d_item_param" % epoch] = model.embedding_dict["item_emb"].detach().numpy().copy()
        shell.predict = lambda u, ue=ue, ie=ie: torch.matmul(ue[data.get_user_id(u)], ie.transpose(0, 1)).numpy()  # LightGCN.py:99-102
        rec_list = GraphRecommender.test(shell)
        out["epoch%d_measures" % epoch] = np.array(ranking_evaluation(data.test_set, rec_list, TOPN))
    out["batch_sizes"] = np.array([t.shape[1] for t in triples])
    out["triples"] = np.concatenate(triples, axis=1)
    out["losses"] = np.array(losses, dtype=np.float64)
    out["batch_epoch"] = np.array(batch_epoch)
    np.savez_compressed(os.path.join(HERE, "lightgcn_loop.npz"), **out)
    print("wrote lightgcn_loop.npz: %d batches, losses %s ... %s" % (len(losses), losses[0], losses[-1]))
    print("".join(out["epoch%d_measures" % (EPOCHS - 1)]))


if __name__ == "__main__":
    main()
