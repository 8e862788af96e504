#!/usr/bin/env python
"""Golden vectors for two more callers of the hot path, produced by the UNMODIFIED reference classes on the CPU:
``SHTEncoder`` (model/graph/SHT.py:142-203: LightGCN-sum propagation + low-rank hypergraph transform) and ``DHCF_Encoder``
(model/graph/DHCF.py:135-186: HGCNConv on the RECTANGULAR user x item interaction matrix).

    python tests/golden/make_golden_more_encoders.py      # needs /root/reference; rewrites tests/golden/more_encoders.npz

Stored: the interaction lists, seeded parameters, forward outputs, and the gradients of a fixed scalar functional of the outputs.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

REF = os.environ.get("HGR_REFERENCE", "/root/reference/HD_SELFRec")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")


def main():
    import torch

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.set_num_threads(1)
    os.chdir(tempfile.mkdtemp(prefix="hgr_golden_enc_"))
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from data.ui_graph import Interaction
    from model.graph.DHCF import DHCF_Encoder
    from model.graph.SHT import SHTEncoder

    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(70, 110, 900, seed=13)
    train = [[int(u), int(i) + 1000, 1.0] for u, i in zip(g.train_u, g.train_i)]
    test = [[int(u), int(i) + 1000, 1.0] for u, i in zip(g.test_u, g.test_i)]
    data = Interaction(None, [list(e) for e in train], [list(e) for e in test])
    out = {"train": np.array(train), "test": np.array(test)}
    rng = np.random.default_rng(3)

    # ---- SHT
    args = {"max_epoch": 1, "batch_size": 64, "lrate": 0.01, "lr_decay": 0.9, "reg": 0.0, "embedding_size": 64, "hyper_dim": 64,
            "drop_rate": 0.2, "p": 0.1, "n_layers": 2, "cl_rate": 1e-4, "temp": 0.2, "seed": 20, "early_stopping_steps": 20,
            "hyperedge_num": 32}
    torch.manual_seed(20)
    sht = SHTEncoder(data, args)
    for k, v in sht.state_dict().items():
        out["sht_param/" + k] = v.detach().numpy().copy()
    emb, hu, hi = sht()
    w = [torch.from_numpy(rng.standard_normal(t.shape).astype(np.float32)) for t in (emb, hu, hi)]
    (emb * w[0]).sum().add((hu * w[1]).sum()).add((hi * w[2]).sum()).backward()
    out["sht_embeds"], out["sht_hyper_u"], out["sht_hyper_i"] = (t.detach().numpy().copy() for t in (emb, hu, hi))
    out["sht_w0"], out["sht_w1"], out["sht_w2"] = (t.numpy() for t in w)
    for k, p in sht.named_parameters():
        out["sht_grad/" + k] = p.grad.numpy().copy()
    out["sht_args"] = np.array([args["n_layers"], args["hyper_dim"], args["hyperedge_num"]])

    # ---- DHCF
    dargs = {"input_dim": 64, "hyper_dim": 64, "p": 0.3, "drop_rate": 0.2, "n_layers": 2}
    torch.manual_seed(21)
    dh = DHCF_Encoder(None, data, dargs)
    for k, v in dh.state_dict().items():
        out["dhcf_param/" + k] = v.detach().numpy().copy()
    ue, ie = dh()
    wu = torch.from_numpy(rng.standard_normal(ue.shape).astype(np.float32))
    wi = torch.from_numpy(rng.standard_normal(ie.shape).astype(np.float32))
    ((ue * wu).sum() + (ie * wi).sum()).backward()
    out["dhcf_user_out"], out["dhcf_item_out"], out["dhcf_wu"], out["dhcf_wi"] = ue.detach().numpy().copy(), ie.detach().numpy().copy(), wu.numpy(), wi.numpy()
    out["dhcf_grad_user"] = dh.embedding_dict["user_emb"].grad.numpy().copy()
    out["dhcf_grad_item"] = dh.embedding_dict["item_emb"].grad.numpy().copy()
    out["dhcf_args"] = np.array([dargs["n_layers"], dargs["hyper_dim"], dargs["p"]])
    np.savez_compressed(os.path.join(HERE, "more_encoders.npz"), **out)
    print("wrote more_encoders.npz", {k: v.shape for k, v in out.items() if "param" in k})


if __name__ == "__main__":
    main()
