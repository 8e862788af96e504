"""Two-GPU tests of the row-partitioned path with the FUSED all-gather (propagation epilogue storing into every
rank's gathered table through symmetric memory).  Needs >= 2 CUDA devices on one box; skipped otherwise
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_USERS, N_ITEMS, D = 3001, 4999, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def _worker(rank, world, port):
    import torch.distributed as dist
    import torch.nn.functional as F

    from hypergraph_diffusion_for_recommendation_b200 import dist as hdist
    from hypergraph_diffusion_for_recommendation_b200 import encoders, graph, ops
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = powerlaw_interactions(N_USERS, N_ITEMS, 120_000, seed=5)
        u, i = torch.from_numpy(g.train_u).to(dev), torch.from_numpy(g.train_i).to(dev)
        ctx = hdist.build_partitioned(u, i, N_USERS, N_ITEMS, rank, world, dev)
        adj, part = ctx.adj, ctx.part
        assert adj.fused, "fused all-gather should be on for a CUDA DistGraph"
        os.environ["HGR_FUSED_GATHER"] = "0"
        plain = hdist.build_partitioned(u, i, N_USERS, N_ITEMS, rank, world, dev).adj
        os.environ["HGR_FUSED_GATHER"] = "1"
        assert not plain.fused
        whole = graph.build_norm_adj(u, i, N_USERS, N_ITEMS, device=dev)  # the unsharded matrix on this GPU
        perm = torch.cat([part.perm_user(torch.arange(N_USERS, device=dev)), part.perm_item(torch.arange(N_ITEMS, device=dev))])
        live = torch.zeros(part.n_glob, dtype=torch.bool, device=dev)
        live[perm] = True
        live = live[rank * part.n_loc:(rank + 1) * part.n_loc]

        def own(t):
            full = torch.zeros(part.n_glob, t.shape[1], device=dev)
            full[perm] = t
            return full[rank * part.n_loc:(rank + 1) * part.n_loc].clone()

        torch.manual_seed(7)
        e_glob = torch.randn(N_USERS + N_ITEMS, D, device=dev) * 0.1
        g_glob = torch.randn(N_USERS + N_ITEMS, D, device=dev)

        # ---- LightGCN: fused == unfused == unsharded (bit for bit: the row accumulation order never changes)
        x1 = own(e_glob).requires_grad_(True)
        o1 = adj.lightgcn_propagate(x1, 3)
        x2 = own(e_glob).requires_grad_(True)
        o2 = plain.lightgcn_propagate(x2, 3)
        xw = e_glob.clone().requires_grad_(True)
        ow = ops.lightgcn_propagate(whole, xw, 3)
        assert torch.equal(o1[live], o2[live]) and torch.equal(o1[live], own(ow.detach())[live])
        (o1 * own(g_glob)).sum().backward()
        (o2 * own(g_glob)).sum().backward()
        (ow * g_glob).sum().backward()
        assert torch.equal(x1.grad[live], x2.grad[live]) and torch.equal(x1.grad[live], own(xw.grad)[live])
        assert adj.n_fused >= 4 and plain.n_fused == 0

        # ---- two chained hypergraph convolutions with LayerNorm + residual (EquivSetConv's pattern): the second one
        # finds its input already gathered (published by the first one's node stage)
        gamma = torch.randn(D, device=dev)
        beta = torch.randn(D, device=dev)

        def chain(a, x, gm, bt):
            h = ops.hgconv(a, x, 0.5, gm, bt, residual=x)
            return ops.hgconv(a, h, 0.5, gm, bt, residual=h)

        outs = []
        for a, xin in ((adj, own(e_glob)), (plain, own(e_glob)), (whole, e_glob.clone())):
            xin = xin.requires_grad_(True)
            gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
            before, before_pub = getattr(a, "n_collective", 0), getattr(a, "n_published", 0)
            y = chain(a, xin, gm, bt)
            w = g_glob if a is whole else own(g_glob)
            m = torch.ones_like(y[:, :1]) if a is whole else live[:, None].float()
            (y * w * m).sum().backward()
            colls = getattr(a, "n_collective", 0) - before  # forward + backward
            if a is not whole:
                dist.all_reduce(gm.grad)
                dist.all_reduce(bt.grad)
            outs.append((y.detach(), xin.grad, gm.grad, bt.grad, (colls, getattr(a, "n_published", 0) - before_pub)))
        (y1, dx1, dg1, db1, c1), (y2, dx2, dg2, db2, c2), (yw, dxw, dgw, dbw, _) = outs
        # NCCL-only: 2 gathers per convolution and direction.  Fused: none at all -- the first input is published by the copy
        # kernel, every other table by the kernel that computes its rows (propagation epilogue, LayerNorm backward)
        assert c1 == (0, 1) and c2 == (8, 0), (c1, c2)
        assert torch.equal(y1[live], y2[live]) and torch.equal(y1[live], own(yw)[live])
        assert torch.equal(dx1[live], dx2[live]) and _rel(dx1[live], own(dxw)[live]) < 1e-5
        assert _rel(dg1, dgw) < 1e-4 and _rel(db1, dbw) < 1e-4

        # ---- one sharded training step of the HGNN_HD3 local encoder, fused vs unfused
        losses = []
        for a in (adj, plain):
            data = type("D", (), {})()
            data.n_users, data.n_items, data.norm_adj, data.norm_adj_device = part.up, part.n_loc - part.up, None, a
            torch.manual_seed(11)
            model = encoders.HGNNModel(data, {"hyper_dim": D, "n_layers": 2}).to(dev)
            model.eval()
            opt = torch.optim.Adam(model.parameters(), lr=1e-3)
            gen = torch.Generator(device=dev)
            gen.manual_seed(3)
            pick = torch.randint(0, u.numel(), (4096,), device=dev, generator=gen)
            neg = torch.randint(0, N_ITEMS, (4096,), device=dev, generator=gen)
            out = hdist.train_step(model, opt, a, u[pick], i[pick], neg, 0.01, 4096)
            losses.append((out.clone(), model.embedding_dict["user_emb"].detach().clone()))
        # (the loss backward scatters with float atomics, so the updated rows agree to rounding, not to the bit)
        assert torch.equal(losses[0][0], losses[1][0]) and _rel(losses[0][1], losses[1][1]) < 1e-5
        both = torch.stack([losses[0][0]])
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        assert torch.equal(both[0], losses[0][0])  # every rank computed the same loss
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_fused_all_gather_matches_nccl_and_unsharded():
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
