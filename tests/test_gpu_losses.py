"""GPU parity of the fused loss kernels against the reference's torch losses (golden vectors) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import hgr_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # north_star tolerance for losses in fp32


from conftest import global_rel_err as grad_err  # noqa: E402  gradients: relative to the largest row (see conftest)
from conftest import rowwise_rel_err as rel_err  # noqa: E402  losses / forward values


@pytest.fixture(scope="module")
def L():
    from hypergraph_diffusion_for_recommendation_b200 import loss_torch

    return loss_torch


def test_bpr_l2_matches_reference_losses_and_gradients(L, golden):
    ut = torch.from_numpy(golden["loss_user_tab"]).cuda().requires_grad_(True)
    it = torch.from_numpy(golden["loss_item_tab"]).cuda().requires_grad_(True)
    # the reference's sampler hands over CPU LongTensors (util/sampler.py:261-263)
    u, p, n = (torch.from_numpy(golden[k]) for k in ("tri_u", "tri_p", "tri_n"))
    rec, reg = L.bpr_l2_from_tables(ut, it, u, p, n, float(golden["loss_reg_lambda"]), int(golden["loss_reg_batch_size"]))
    assert rel_err(rec, golden["loss_bpr"]) < RTOL and rel_err(reg, golden["loss_reg"]) < RTOL
    (rec + reg).backward()
    assert grad_err(ut.grad, golden["loss_dU"]) < RTOL and grad_err(it.grad, golden["loss_dI"]) < RTOL
    o_rec, o_reg, o_du, o_di = O.bpr_l2_from_tables(golden["loss_user_tab"], golden["loss_item_tab"], golden["tri_u"], golden["tri_p"],
                                                    golden["tri_n"], float(golden["loss_reg_lambda"]), int(golden["loss_reg_batch_size"]))
    assert rel_err(rec, o_rec) < RTOL and grad_err(ut.grad, o_du) < RTOL and grad_err(it.grad, o_di) < RTOL


@pytest.mark.parametrize("d,batch", [(32, 1), (64, 4097), (128, 300)])
def test_bpr_l2_shapes_against_oracle(L, d, batch):
    rng = np.random.default_rng(batch)
    ut = rng.standard_normal((50, d)).astype(np.float32)
    it = rng.standard_normal((80, d)).astype(np.float32) * 0.5
    u, p, n = rng.integers(0, 50, batch), rng.integers(0, 80, batch), rng.integers(0, 80, batch)
    tu, ti = torch.from_numpy(ut).cuda().requires_grad_(True), torch.from_numpy(it).cuda().requires_grad_(True)
    rec, reg = L.bpr_l2_from_tables(tu, ti, torch.from_numpy(u).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(n).cuda(), 0.05, 2048)
    (2.0 * rec + 3.0 * reg).backward()
    o_rec, o_reg, o_du, o_di = O.bpr_l2_from_tables(ut, it, u, p, n, 0.05, 2048)
    assert rel_err(rec, o_rec) < RTOL and rel_err(reg, o_reg) < RTOL
    # the oracle returns d(rec + reg); rebuild the weighted gradient from two oracle calls
    _, _, du0, di0 = O.bpr_l2_from_tables(ut, it, u, p, n, 0.0, 2048)  # d rec
    want_du, want_di = 2.0 * du0 + 3.0 * (o_du - du0), 2.0 * di0 + 3.0 * (o_di - di0)
    assert grad_err(tu.grad, want_du) < 2e-5 and grad_err(ti.grad, want_di) < 2e-5


@pytest.mark.parametrize("world,d,batch", [(2, 64, 5000), (4, 32, 777), (3, 128, 64)])
def test_bpr_l2_owner_sharded_equals_the_whole_batch(L, world, d, batch):
    """dist.train_step's loss at world > 1, replayed on one GPU: every "rank" evaluates the triples of the users in its window of
    the gathered table (hgr_bpr_l2_fwd_owned_f32), the four sums are added (the all_reduce), and each rank's backward writes its
    own rows (hgr_bpr_l2_bwd_owned_f32).  Against the oracle on the whole batch: losses to 1e-5, the stacked gradient rows to
    1e-5 of the largest; the loss is the same on every rank."""
    rng = np.random.default_rng(world * 1000 + d)
    n_loc = 96
    tab = (rng.standard_normal((world * n_loc, d)) * 0.3).astype(np.float32)
    u = rng.integers(0, world * n_loc, batch)
    p = rng.integers(0, world * n_loc, batch)
    n = rng.integers(0, world * n_loc, batch)
    tu, tp, tn = (torch.from_numpy(v).cuda() for v in (u, p, n))
    full = torch.from_numpy(tab).cuda()
    # pass 1: every rank's partial sums (what crosses the ranks)
    parts = []

    def grab(t):
        parts.append(t.clone())

    for r in range(world):
        own = full[r * n_loc:(r + 1) * n_loc].clone().requires_grad_(True)
        L.bpr_l2_sharded(own, lambda x: full, r * n_loc, tu, tp, tn, 0.05, 2048, grab)
    total = torch.stack(parts).sum(0)
    assert all(float(q[0]) > 0 for q in parts)  # every rank owned some triples
    # pass 2: the step itself, the all_reduce replaced by the precomputed total
    losses, grads = [], []
    for r in range(world):
        own = full[r * n_loc:(r + 1) * n_loc].clone().requires_grad_(True)
        rec, reg = L.bpr_l2_sharded(own, lambda x: full, r * n_loc, tu, tp, tn, 0.05, 2048, lambda t: t.copy_(total))
        (2.0 * rec + 3.0 * reg).backward()
        losses.append(torch.stack([rec.detach(), reg.detach()]))
        grads.append(own.grad)
    for q in losses[1:]:
        assert torch.equal(q, losses[0])
    o_rec, o_reg, o_du, o_di = O.bpr_l2_from_tables(tab, tab, u, p, n, 0.05, 2048)
    _, _, du0, di0 = O.bpr_l2_from_tables(tab, tab, u, p, n, 0.0, 2048)
    want = 2.0 * (du0 + di0) + 3.0 * ((o_du - du0) + (o_di - di0))  # users and items index the same table
    assert rel_err(losses[0][0], o_rec) < RTOL and rel_err(losses[0][1], o_reg) < RTOL
    assert grad_err(torch.cat(grads), want) < 2e-5
    # and the replicated form (every rank evaluates everything) gives the same numbers
    own = full[:n_loc].clone().requires_grad_(True)
    rec2, reg2 = L.bpr_l2_sharded(own, lambda x: full, 0, tu, tp, tn, 0.05, 2048)
    (2.0 * rec2 + 3.0 * reg2).backward()
    assert rel_err(rec2, losses[0][0]) < 1e-6 and rel_err(reg2, losses[0][1]) < 1e-6
    assert grad_err(own.grad, grads[0]) < 2e-5


def test_reference_signature_bpr_loss_and_bad_indices(L, golden):
    ut, it = golden["loss_user_tab"], golden["loss_item_tab"]
    u, p, n = golden["tri_u"], golden["tri_p"], golden["tri_n"]
    ue, pe, ne = (torch.from_numpy(a).cuda() for a in (ut[u], it[p], it[n]))
    assert rel_err(L.bpr_loss(ue, pe, ne), golden["loss_bpr"]) < RTOL
    assert rel_err(L.l2_reg_loss(0.1, ue, pe, ne) / 2048, golden["loss_reg"]) < RTOL
    with pytest.raises(Exception):
        L.bpr_l2_from_tables(torch.from_numpy(ut), torch.from_numpy(it), torch.from_numpy(u), torch.from_numpy(p), torch.from_numpy(n), 0.1, 2048)


# ------------------------------------------------------------------------------------ contrastLoss / InfoNCE (csrc/loss_ssl.cu)
def test_contrast_loss_and_infonce_match_reference_golden(L, golden):
    e1 = torch.from_numpy(golden["cl_e1"]).cuda().requires_grad_(True)
    e2 = torch.from_numpy(golden["cl_e2"]).cuda().requires_grad_(True)
    loss = L.contrastLoss(e1, e2, torch.from_numpy(golden["cl_nodes"]), float(golden["cl_temp"]))
    assert rel_err(loss, golden["cl_loss"]) < RTOL
    loss.backward()
    assert grad_err(e1.grad, golden["cl_d1"]) < 5e-5 and grad_err(e2.grad, golden["cl_d2"]) < 5e-5
    v1 = torch.from_numpy(golden["nce_v1"]).cuda().requires_grad_(True)
    v2 = torch.from_numpy(golden["nce_v2"]).cuda().requires_grad_(True)
    loss = L.InfoNCE(v1, v2, float(golden["nce_temp"]))
    assert rel_err(loss, golden["nce_loss"]) < RTOL
    loss.backward()
    assert grad_err(v1.grad, golden["nce_d1"]) < 5e-5 and grad_err(v2.grad, golden["nce_d2"]) < 5e-5


@pytest.mark.parametrize("d,n_rows,m", [(32, 40, 1), (64, 3000, 1111), (64, 5000, 4096), (128, 700, 65)])
def test_contrast_loss_shapes_against_oracle_and_torch(L, d, n_rows, m):
    import torch.nn.functional as F

    rng = np.random.default_rng(m)
    e1 = rng.standard_normal((n_rows, d)).astype(np.float32) * 0.3
    e2 = rng.standard_normal((n_rows, d)).astype(np.float32) * 0.3
    nodes = np.sort(rng.permutation(n_rows)[:m])
    t1, t2 = torch.from_numpy(e1).cuda().requires_grad_(True), torch.from_numpy(e2).cuda().requires_grad_(True)
    loss = L.contrastLoss(t1, t2, torch.from_numpy(nodes).cuda(), 0.2)
    (3.0 * loss).backward()
    o_loss, o_d1, o_d2 = O.contrast_loss(e1, e2, nodes, 0.2)
    if m == 1:  # a single node: loss = log(1 + 1e-8 / e^x) ~ 0, gradients ~ 0 (pure cancellation)
        assert abs(float(loss.detach()) - float(o_loss)) < 1e-6 and float(t1.grad.abs().max()) < 1e-6
        return
    assert rel_err(loss, o_loss) < RTOL
    assert grad_err(t1.grad, 3.0 * o_d1) < 5e-5 and grad_err(t2.grad, 3.0 * o_d2) < 5e-5
    # the reference's own expression in torch on the GPU (util/loss_torch.py:103-110), HCCF-style detached first view
    r1, r2 = torch.from_numpy(e1).cuda(), torch.from_numpy(e2).cuda().requires_grad_(True)
    a, b = F.normalize(r1 + 1e-8, p=2)[nodes], F.normalize(r2 + 1e-8, p=2)[nodes]
    ref = -torch.log(torch.exp((a * b).sum(-1) / 0.2) / (torch.exp(a @ b.T / 0.2).sum(-1) + 1e-8)).mean()
    ref.backward()
    d2 = torch.from_numpy(e2).cuda().requires_grad_(True)
    got = L.contrastLoss(torch.from_numpy(e1).cuda(), d2, torch.from_numpy(nodes), 0.2)  # CPU index tensor, detached e1
    got.backward()
    assert rel_err(got, ref) < RTOL and grad_err(d2.grad, r2.grad) < 5e-5


@pytest.mark.parametrize("b_cos", [True, False])
def test_infonce_against_torch_expression(L, b_cos):
    import torch.nn.functional as F

    rng = np.random.default_rng(17)
    v1 = (rng.standard_normal((777, 64)) * (0.2 if not b_cos else 1.0)).astype(np.float32)
    v2 = (v1 + 0.1 * rng.standard_normal((777, 64)) * (0.2 if not b_cos else 1.0)).astype(np.float32)
    t1, t2 = torch.from_numpy(v1).cuda().requires_grad_(True), torch.from_numpy(v2).cuda().requires_grad_(True)
    loss = L.InfoNCE(t1, t2, 0.2, b_cos)
    loss.backward()
    r1, r2 = torch.from_numpy(v1).cuda().requires_grad_(True), torch.from_numpy(v2).cuda().requires_grad_(True)
    a, b = (F.normalize(r1, dim=1), F.normalize(r2, dim=1)) if b_cos else (r1, r2)
    pos = torch.exp((a * b).sum(-1) / 0.2)
    ttl = torch.exp(a @ b.T / 0.2).sum(1)
    ref = (-torch.log(pos / ttl + 10e-6)).mean()
    ref.backward()
    assert rel_err(loss, ref) < RTOL and grad_err(t1.grad, r1.grad) < 5e-5 and grad_err(t2.grad, r2.grad) < 5e-5
    if b_cos:
        o_loss, o_d1, o_d2 = O.info_nce(v1, v2, 0.2)
        assert rel_err(loss, o_loss) < RTOL and grad_err(t1.grad, o_d1) < 5e-5


@pytest.mark.parametrize("batch,n_rows", [(1, 50), (700, 300), (4096, 3000)])
def test_contrast_loss_padded_batch_equals_unique(L, batch, n_rows):
    """``contrastLoss_padded(e1, e2, batch_ids)`` -- the fixed-shape, CUDA-graph-capturable form (sorted ids, repeats replaced
    by inactive slots) -- against ``contrastLoss(e1, e2, torch.unique(batch_ids))`` and the oracle: loss and both gradients."""
    rng = np.random.default_rng(batch)
    e1 = rng.standard_normal((n_rows, 64)).astype(np.float32) * 0.3
    e2 = rng.standard_normal((n_rows, 64)).astype(np.float32) * 0.3
    ids = rng.integers(0, n_rows, batch)  # with repeats
    t_ids = torch.from_numpy(ids).cuda()
    pad = L.unique_padded(t_ids)
    uniq = np.unique(ids)
    assert pad.shape == t_ids.shape and np.array_equal(pad[pad >= 0].cpu().numpy(), uniq)  # torch.unique's order, gaps marked -1
    a1, a2 = torch.from_numpy(e1).cuda().requires_grad_(True), torch.from_numpy(e2).cuda().requires_grad_(True)
    b1, b2 = torch.from_numpy(e1).cuda().requires_grad_(True), torch.from_numpy(e2).cuda().requires_grad_(True)
    lp = L.contrastLoss_padded(a1, a2, t_ids, 0.2)
    lu = L.contrastLoss(b1, b2, torch.unique(t_ids), 0.2)
    (2.0 * lp).backward()
    (2.0 * lu).backward()
    if uniq.size == 1:
        assert abs(float(lp) - float(lu)) < 1e-6
        return
    assert rel_err(lp, lu) < 1e-6  # same terms, tiles cut differently
    assert grad_err(a1.grad, b1.grad) < 1e-5 and grad_err(a2.grad, b2.grad) < 1e-5
    o_loss, o_d1, o_d2 = O.contrast_loss(e1, e2, uniq, 0.2)
    assert rel_err(lp, o_loss) < RTOL and grad_err(a1.grad, 2.0 * o_d1) < 5e-5 and grad_err(a2.grad, 2.0 * o_d2) < 5e-5
    touched = np.zeros(n_rows, dtype=bool)
    touched[uniq] = True
    assert not a1.grad[torch.from_numpy(~touched).cuda()].any()  # rows outside the batch get no gradient


def test_hccf_step_replays_as_a_cuda_graph():
    """The HCCF step with ``static_shapes=True`` captured once and replayed: losses equal to the eager step on the same
    batch and random streams are finite and move the parameters (no torch.unique, no host sync inside the step)."""
    from hypergraph_diffusion_for_recommendation_b200 import encoders, graph, trainer
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(400, 600, 9000, seed=2)
    dev = torch.device("cuda")
    u, i = torch.from_numpy(g.train_u).to(dev), torch.from_numpy(g.train_i).to(dev)
    data = type("D", (), {})()
    data.n_users, data.n_items = 400, 600
    data.norm_adj = graph.build_norm_adj(u, i, 400, 600, device=dev)
    conf = {"lrate": 0.001, "lr_decay": 1.0, "max_epoch": 1, "batch_size": 512, "reg": 0.0, "embedding_size": 64, "hyper_dim": 128,
            "drop_rate": 0.0, "p": 0.5, "n_layers": 2}
    torch.manual_seed(5)
    model = encoders.HCCFEncoder(conf, data).to(dev)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=0.001, fused=True, capturable=True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    pick = torch.randint(0, u.numel(), (512,), device=dev, generator=gen)
    tu, tp, tn = u[pick].long(), i[pick].long(), torch.randint(0, 600, (512,), device=dev, generator=gen)
    # keep_rate 1 / drop_rate 0: no random stream inside the step, so eager and replayed losses can be compared
    eager = trainer.train_step_hccf(model, opt, tu, tp, tn, 0.2, 0.1, 1.0, static_shapes=False).clone()
    torch.manual_seed(5)
    model2 = encoders.HCCFEncoder(conf, data).to(dev)
    model2.train()
    opt2 = torch.optim.Adam(model2.parameters(), lr=0.001, fused=True, capturable=True)
    before = {k: v.detach().clone() for k, v in model2.state_dict().items()}
    step = trainer.GraphedStep(lambda a, b, c: trainer.train_step_hccf(model2, opt2, a, b, c, 0.2, 0.1, 1.0, static_shapes=True), 512, dev,
                               warmup=1)
    # the warm-up step ran on zero indices: put parameters and Adam state back IN PLACE (the graph holds their addresses)
    model2.load_state_dict(before)
    for st in opt2.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    out = step(tu, tp, tn).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and step.kernels_per_replay > 20
    assert rel_err(out, eager) < 1e-5
    assert (model2.embedding_dict["user_emb"] - before["embedding_dict.user_emb"]).abs().max() > 0
