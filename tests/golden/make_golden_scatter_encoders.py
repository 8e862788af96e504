#!/usr/bin/env python
"""Golden vectors for the consumers of the scatter form, produced by the UNMODIFIED reference classes on the CPU:

* ``HGNN_HD4.LocalAwareEncoder`` (model/graph/HGNN_HD4.py:337-405) on ``EquivSetGNN2`` / ``EquivSetConv2``
  (model/layers/layers2/EquivSetGNN2.py:83-133, EquivSetConv2.py:85-100): star expansion of the dense (U+I)^2 adjacency;
* ``HCCF_diffusion.HCCFEncoder`` (model/graph/HCCF_diffusion.py:131-217,291-308,382-402): sign-thresholded learned incidence;
* ``HD2.EquivSetConv`` (model/graph/HD2.py:589-643): the per-edge attention-weighted variant (width 32 is hard-coded there);
* ``Graph.normalize_graph_mat_hyper`` (data/graph.py:28-42).

    python tests/golden/make_golden_scatter_encoders.py   # needs /root/reference; rewrites tests/golden/scatter_encoders.npz

``torch_scatter`` is absent here: the stub of make_golden.py stands in (pytorch-scatter 2.1.0 semantics: sum via index_add,
mean = sum / clamp(count, 1)).  All modules run in eval mode (dropout = identity), edge keep rate 1.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

REF = os.environ.get("HGR_REFERENCE", "/root/reference/HD_SELFRec")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")


def main():
    import torch
    from make_golden import install_shims

    install_shims()
    torch.set_num_threads(1)
    os.chdir(tempfile.mkdtemp(prefix="hgr_golden_scat_"))
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from data.graph import Graph
    from data.ui_graph import Interaction

    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(70, 110, 900, seed=13)
    train = [[int(u), int(i) + 1000, 1.0] for u, i in zip(g.train_u, g.train_i)]
    test = [[int(u), int(i) + 1000, 1.0] for u, i in zip(g.test_u, g.test_i)]
    data = Interaction(None, [list(e) for e in train], [list(e) for e in test])
    out = {"train": np.array(train), "test": np.array(test)}
    rng = np.random.default_rng(5)
    n = data.n_users + data.n_items

    # ---- normalize_graph_mat_hyper on the interaction matrix (users x items incidence) and on the bipartite adjacency
    for name, mat in (("inter", data.interaction_mat), ("adj", data.ui_adj)):
        m = Graph.normalize_graph_mat_hyper(mat).tocsr()
        m.sort_indices()
        out["hyper_%s_indptr" % name], out["hyper_%s_indices" % name], out["hyper_%s_data" % name] = m.indptr, m.indices, m.data.astype(np.float32)
        x = rng.standard_normal((mat.shape[0], 64)).astype(np.float32)
        out["hyper_%s_X" % name] = x
        out["hyper_%s_Y" % name] = (m @ x).astype(np.float32)

    # ---- HGNN_HD4 LocalAwareEncoder
    import model.graph.HGNN_HD4 as HD4

    torch.manual_seed(31)
    lae = HD4.LocalAwareEncoder(data, 64, 64, 2, 0.3, 0.2, torch.device("cpu"))
    lae.eval()
    for k, v in lae.state_dict().items():
        out["hd4_param/" + k] = v.detach().numpy().copy()
    e0 = torch.from_numpy((rng.standard_normal((n, 64)) * 0.1).astype(np.float32)).requires_grad_(True)
    ue, ie = lae(e0, lae.sparse_norm_adj)
    w = torch.from_numpy(rng.standard_normal((n, 64)).astype(np.float32))
    (torch.cat([ue, ie], 0) * w).sum().backward()
    out["hd4_E0"], out["hd4_W"] = e0.detach().numpy().copy(), w.numpy()
    out["hd4_user_out"], out["hd4_item_out"] = ue.detach().numpy().copy(), ie.detach().numpy().copy()
    out["hd4_dE0"] = e0.grad.numpy().copy()
    for k, p in lae.named_parameters():
        if p.grad is not None:
            out["hd4_grad/" + k] = p.grad.numpy().copy()

    # ---- HCCF_diffusion encoder
    import model.graph.HCCF_diffusion as HDF

    conf = dict(lrate=0.001, lr_decay=0.9, max_epoch=1, batch_size=64, reg=0.1, embedding_size=64, hyper_dim=32, drop_rate=0.5, p=0.1,
                n_layers=2)
    torch.manual_seed(32)
    enc = HDF.HCCFEncoder(conf, data)
    enc.eval()
    for k, v in enc.state_dict().items():
        out["hdf_param/" + k] = v.detach().numpy().copy()
    hu, hi, gcn_h, hyp_h = enc(keep_rate=1.0)
    wu = torch.from_numpy(rng.standard_normal(hu.shape).astype(np.float32))
    wi = torch.from_numpy(rng.standard_normal(hi.shape).astype(np.float32))
    ((hu * wu).sum() + (hi * wi).sum()).backward()
    out["hdf_user_out"], out["hdf_item_out"], out["hdf_wu"], out["hdf_wi"] = hu.detach().numpy().copy(), hi.detach().numpy().copy(), wu.numpy(), wi.numpy()
    for l in range(2):
        out["hdf_gcn_%d" % l], out["hdf_hyp_%d" % l] = gcn_h[l].detach().numpy().copy(), hyp_h[l].detach().numpy().copy()
    for k, p in enc.named_parameters():
        out["hdf_grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    out["hdf_conf"] = np.array([conf["n_layers"], conf["hyper_dim"], conf["embedding_size"]])

    # ---- HD2: attention-weighted scatter convolution (W1 = identity, W2 = slice, W = Linear(LN(.)), alpha = 0)
    import model.graph.HD2 as HD2

    torch.manual_seed(33)
    conv = HD2.EquivSetConv(32, 32, mlp1_layers=0, mlp2_layers=0, mlp3_layers=1, alpha=0.0, aggr="mean", dropout=0.5, normalization="ln",
                            input_norm=True)
    conv.eval()
    for k, v in conv.state_dict().items():
        out["att_param/" + k] = v.detach().numpy().copy()
    n_nodes, n_edges, nnz = 90, 17, 400
    pairs = np.unique(np.stack([rng.integers(0, n_nodes, nnz), rng.integers(0, n_edges, nnz)], 1), axis=0)
    V, E = torch.from_numpy(pairs[:, 0]), torch.from_numpy(pairs[:, 1])
    x = torch.from_numpy(rng.standard_normal((n_nodes, 32)).astype(np.float32)).requires_grad_(True)
    atts = torch.from_numpy(rng.random((pairs.shape[0], 1)).astype(np.float32)).requires_grad_(True)
    y = conv(x, V, E, atts, x)
    wy = torch.from_numpy(rng.standard_normal(y.shape).astype(np.float32))
    (y * wy).sum().backward()
    out["att_V"], out["att_E"], out["att_X"], out["att_atts"], out["att_Y"], out["att_W"] = pairs[:, 0], pairs[:, 1], x.detach().numpy().copy(), atts.detach().numpy().copy(), y.detach().numpy().copy(), wy.numpy()
    out["att_dX"], out["att_datts"] = x.grad.numpy().copy(), atts.grad.numpy().copy()
    for k, p in conv.named_parameters():
        out["att_grad/" + k] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "scatter_encoders.npz"), **out)
    print("wrote scatter_encoders.npz:", sorted(k for k in out if "param" in k))


if __name__ == "__main__":
    main()
