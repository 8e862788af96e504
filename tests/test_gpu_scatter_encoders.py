"""GPU parity of the consumers of the scatter form against vectors produced by the UNMODIFIED reference classes
(tests/golden/scatter_encoders.npz, make_golden_scatter_encoders.py): HGNN_HD4's LocalAwareEncoder, HCCF_diffusion's encoder,
HD2's attention-weighted convolution, Graph.normalize_graph_mat_hyper.  Forward 1e-5 row-wise (north_star); gradients relative to
the largest row (conftest.global_rel_err); checkpoints load with strict=True (same state_dict keys as the reference)."""
import os

import numpy as np
import pytest
import torch
from conftest import global_rel_err as grad_err
from conftest import rowwise_rel_err as rel_err

from oracle import hgr_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def scat():
    return np.load(os.path.join(HERE, "golden", "scatter_encoders.npz"))


@pytest.fixture(scope="module")
def data(scat):
    from hypergraph_diffusion_for_recommendation_b200 import data as D

    return D.Interaction(None, scat["train"].tolist(), scat["test"].tolist())


def params(npz, prefix):
    return {k[len(prefix):]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(prefix)}


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_hd4_local_aware_encoder_matches_the_reference(scat, data):
    from hypergraph_diffusion_for_recommendation_b200.encoders_scatter import LocalAwareEncoderHD4

    enc = LocalAwareEncoderHD4(data, 64, 64, 2, 0.3, 0.2)
    enc.load_state_dict(params(scat, "hd4_param/"), strict=True)
    enc = enc.cuda().eval()
    e0 = cuda(scat["hd4_E0"]).requires_grad_(True)
    ue, ie = enc(e0, enc.sparse_norm_adj)
    assert rel_err(ue, scat["hd4_user_out"]) < RTOL and rel_err(ie, scat["hd4_item_out"]) < RTOL
    (torch.cat([ue, ie], 0) * cuda(scat["hd4_W"])).sum().backward()
    assert grad_err(e0.grad, scat["hd4_dE0"]) < 2e-5
    checked = 0
    for k, p in enc.named_parameters():
        if "hd4_grad/" + k in scat.files:
            assert grad_err(p.grad, scat["hd4_grad/" + k]) < 5e-5, k
            checked += 1
    assert checked >= 8


def test_hccf_diffusion_encoder_matches_the_reference(scat, data):
    from hypergraph_diffusion_for_recommendation_b200.encoders_scatter import HCCFDiffusionEncoder

    nl, hyper, emb = (int(v) for v in scat["hdf_conf"])
    conf = dict(lrate=0.001, lr_decay=0.9, max_epoch=1, batch_size=64, reg=0.1, embedding_size=emb, hyper_dim=hyper, drop_rate=0.5, p=0.1,
                n_layers=nl)
    enc = HCCFDiffusionEncoder(conf, data)
    enc.load_state_dict(params(scat, "hdf_param/"), strict=True)
    enc = enc.cuda().eval()
    hu, hi, gcn_h, hyp_h = enc(keep_rate=1.0)
    assert rel_err(hu, scat["hdf_user_out"]) < RTOL and rel_err(hi, scat["hdf_item_out"]) < RTOL
    for l in range(nl):
        assert rel_err(gcn_h[l], scat["hdf_gcn_%d" % l]) < RTOL and rel_err(hyp_h[l], scat["hdf_hyp_%d" % l]) < RTOL
    ((hu * cuda(scat["hdf_wu"])).sum() + (hi * cuda(scat["hdf_wi"])).sum()).backward()
    for k, p in enc.named_parameters():
        want = scat["hdf_grad/" + k]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if np.abs(want).max() == 0:  # user_w / item_w: the sign pattern passes no gradient, in the reference as here
            assert float(got.abs().max()) == 0.0, k
        else:
            assert grad_err(got, want) < 5e-5, k


def test_attention_weighted_scatter_convolution_matches_the_reference(scat):
    from hypergraph_diffusion_for_recommendation_b200.encoders_scatter import EquivSetConvAttention

    conv = EquivSetConvAttention(32, 32, mlp1_layers=0, mlp2_layers=0, mlp3_layers=1, alpha=0.0, aggr="mean", dropout=0.5,
                                 normalization="ln", input_norm=True)
    conv.load_state_dict(params(scat, "att_param/"), strict=True)
    conv = conv.cuda().eval()
    x = cuda(scat["att_X"]).requires_grad_(True)
    atts = cuda(scat["att_atts"]).requires_grad_(True)
    v, e = cuda(scat["att_V"]), cuda(scat["att_E"])
    y = conv(x, v, e, atts, x)
    assert rel_err(y, scat["att_Y"]) < RTOL
    (y * cuda(scat["att_W"])).sum().backward()
    assert grad_err(x.grad, scat["att_dX"]) < 2e-5 and grad_err(atts.grad, scat["att_datts"]) < 2e-5
    for k, p in conv.named_parameters():
        assert grad_err(p.grad, scat["att_grad/" + k]) < 5e-5, k
    # and against the oracle's restatement on a ragged case: an empty hyperedge, an isolated vertex
    from hypergraph_diffusion_for_recommendation_b200 import graph, ops

    rng = np.random.default_rng(8)
    vv = np.array([0, 0, 1, 3, 3, 3, 5], dtype=np.int64)
    ee = np.array([0, 2, 2, 0, 2, 4, 4], dtype=np.int64)  # hyperedges 1 and 3 are empty; vertices 2 and 4 are isolated
    xx = rng.standard_normal((6, 32)).astype(np.float32)
    aa = rng.random((7, 1)).astype(np.float32)
    inc = graph.build_incidence(cuda(vv), cuda(ee), 6, 5)
    got = ops.scatter_mean_conv_weighted(inc, cuda(xx), cuda(aa), cuda(vv), cuda(ee))
    assert rel_err(got, O.scatter_mean_conv_weighted(vv, ee, xx, aa, 6)) < RTOL


def test_normalize_graph_mat_hyper_factors_are_bit_exact_and_the_operator_matches_the_reference(scat, data):
    from hypergraph_diffusion_for_recommendation_b200 import graph

    for name, mat in (("inter", data.interaction_mat), ("adj", data.ui_adj)):
        op = graph.normalize_graph_mat_hyper(mat)
        ip, ix, dv = mat.to_host()
        (lp, li, lv), (rp, ri, rv) = O.normalize_graph_mat_hyper(ip, ix, dv, mat.shape[1])
        for dev_m, (p_, i_, v_) in ((op.left, (lp, li, lv)), (op.right, (rp, ri, rv))):
            hp, hi_, hv = dev_m.to_host()
            assert np.array_equal(hp, p_) and np.array_equal(hi_, i_)
            assert np.array_equal(hv.view(np.uint32), np.asarray(v_, np.float32).view(np.uint32))  # scipy's roundings, bit for bit
        y = op.matmul(cuda(scat["hyper_%s_X" % name]))
        assert rel_err(y, scat["hyper_%s_Y" % name]) < RTOL  # the reference's materialised matrix applied to the same table
