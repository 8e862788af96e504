"""Self-check of the row-partitioned path on the ranks that are actually running: sharded == NCCL-only == unsharded.

``parity_check`` is called by ``bench.py --gpus N`` (N > 1; the result goes into the JSON line as ``"parity"``) and by
``tests/test_gpu_dist.py``: the single-GPU test box cannot run a 2-rank test, the scaling run can.  It builds a small
power-law graph (3 001 x 4 999 x 120 k) on every rank, partitions it, and compares on this rank's rows

* ``lightgcn_propagate`` forward and backward: fused all-gather (propagation epilogue storing into every rank's table over
  NVLink) vs NCCL ``all_gather`` vs the whole matrix on one GPU -- bit for bit (a row's accumulation order never changes);
* two chained ``hgconv`` with LayerNorm + residual: forward bit for bit, input gradient bit for bit between the two sharded
  forms and to 1e-5 (elementwise, with an absolute floor) against the unsharded one, LayerNorm gradients after all_reduce;
* one sharded training step of the hypergraph-diffusion encoder, fused vs NCCL: identical losses on every rank;
* ``fullrank_topk_sharded`` vs ``evaluation.fullrank_topk`` on the whole tables: ids and scores bit for bit.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

N_USERS, N_ITEMS, N_TRAIN, D = 3001, 4999, 120_000, 64


def elementwise_close(a: torch.Tensor, b: torch.Tensor, rel: float, floor: float = 1e-6) -> bool:
    """|a - b| <= rel * max(|b|, floor * max|b|) for EVERY element."""
    scale = b.abs().clamp(min=float(b.abs().max()) * floor + 1e-30)
    return bool(((a - b).abs() <= rel * scale).all())


def parity_check(rank: int, world: int, dev: torch.device, group=None) -> dict:
    from . import dist as hdist
    from . import encoders, evaluation, graph, ops
    from .synth import powerlaw_interactions

    checks = {}
    g = powerlaw_interactions(N_USERS, N_ITEMS, N_TRAIN, seed=5)
    u, i = torch.from_numpy(g.train_u).to(dev), torch.from_numpy(g.train_i).to(dev)
    prev = os.environ.get("HGR_FUSED_GATHER")
    os.environ["HGR_FUSED_GATHER"] = "1"
    ctx = hdist.build_partitioned(u, i, N_USERS, N_ITEMS, rank, world, dev, group=group)
    adj, part = ctx.adj, ctx.part
    os.environ["HGR_FUSED_GATHER"] = "0"
    plain = hdist.build_partitioned(u, i, N_USERS, N_ITEMS, rank, world, dev, group=group).adj
    if prev is None:
        del os.environ["HGR_FUSED_GATHER"]
    else:
        os.environ["HGR_FUSED_GATHER"] = prev
    checks["fused_path_active"] = bool(adj.fused) and not plain.fused
    whole = graph.build_norm_adj(u, i, N_USERS, N_ITEMS, device=dev)
    perm = torch.cat([part.perm_user(torch.arange(N_USERS, device=dev)), part.perm_item(torch.arange(N_ITEMS, device=dev))])
    live = torch.zeros(part.n_glob, dtype=torch.bool, device=dev)
    live[perm] = True
    live = live[rank * part.n_loc:(rank + 1) * part.n_loc]

    def own(t):
        full = torch.zeros(part.n_glob, t.shape[1], device=dev, dtype=t.dtype)
        full[perm] = t
        return full[rank * part.n_loc:(rank + 1) * part.n_loc].clone()

    torch.manual_seed(7)
    e_glob = torch.randn(N_USERS + N_ITEMS, D, device=dev) * 0.1
    g_glob = torch.randn(N_USERS + N_ITEMS, D, device=dev)

    # ---- LightGCN
    x1 = own(e_glob).requires_grad_(True)
    o1 = adj.lightgcn_propagate(x1, 3)
    x2 = own(e_glob).requires_grad_(True)
    o2 = plain.lightgcn_propagate(x2, 3)
    xw = e_glob.clone().requires_grad_(True)
    ow = ops.lightgcn_propagate(whole, xw, 3)
    checks["lightgcn_fwd_fused_eq_nccl"] = torch.equal(o1[live], o2[live])
    checks["lightgcn_fwd_sharded_eq_unsharded"] = torch.equal(o1[live], own(ow.detach())[live])
    (o1 * own(g_glob)).sum().backward()
    (o2 * own(g_glob)).sum().backward()
    (ow * g_glob).sum().backward()
    checks["lightgcn_bwd_fused_eq_nccl"] = torch.equal(x1.grad[live], x2.grad[live])
    checks["lightgcn_bwd_sharded_eq_unsharded"] = torch.equal(x1.grad[live], own(xw.grad)[live])

    # ---- two chained hypergraph convolutions with LayerNorm + residual
    gamma = torch.randn(D, device=dev)
    beta = torch.randn(D, device=dev)

    def chain(a, x, gm, bt):
        h = ops.hgconv(a, x, 0.5, gm, bt, residual=x)
        return ops.hgconv(a, h, 0.5, gm, bt, residual=h)

    outs = []
    for a, xin in ((adj, own(e_glob)), (plain, own(e_glob)), (whole, e_glob.clone())):
        xin = xin.requires_grad_(True)
        gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        before = (getattr(a, "n_collective", 0), getattr(a, "n_published", 0))
        y = chain(a, xin, gm, bt)
        w = g_glob if a is whole else own(g_glob)
        m = torch.ones_like(y[:, :1]) if a is whole else live[:, None].float()
        (y * w * m).sum().backward()
        counts = (getattr(a, "n_collective", 0) - before[0], getattr(a, "n_published", 0) - before[1])
        if a is not whole:
            dist.all_reduce(gm.grad, group=group)
            dist.all_reduce(bt.grad, group=group)
        outs.append((y.detach(), xin.grad, gm.grad, bt.grad, counts))
    (y1, dx1, dg1, db1, c1), (y2, dx2, dg2, db2, c2), (yw, dxw, dgw, dbw, _) = outs
    checks["hgconv_no_nccl_gather_when_fused"] = c1[0] == 0 and c2 == (8, 0)
    checks["hgconv_fwd_fused_eq_nccl"] = torch.equal(y1[live], y2[live])
    checks["hgconv_fwd_sharded_eq_unsharded"] = torch.equal(y1[live], own(yw)[live])
    checks["hgconv_bwd_fused_eq_nccl"] = torch.equal(dx1[live], dx2[live])
    checks["hgconv_bwd_sharded_vs_unsharded_1e-5"] = elementwise_close(dx1[live], own(dxw)[live], 1e-5, 1e-3)
    checks["layernorm_grads_1e-4"] = elementwise_close(dg1, dgw, 1e-4, 1e-3) and elementwise_close(db1, dbw, 1e-4, 1e-3)

    # ---- one sharded training step, fused vs NCCL
    losses = []
    models = []
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
        models.append(model)
    checks["train_step_loss_fused_eq_nccl"] = torch.equal(losses[0][0], losses[1][0])
    # (the loss backward accumulates with float atomics and the first Adam step turns a gradient into lr * g / (|g| + 1e-8): rows
    # the batch hardly touches move by rounding noise, so the updated table is compared relative to its largest entry)
    diff = float((losses[0][1] - losses[1][1]).abs().max() / losses[1][1].abs().max())
    checks["train_step_rows_fused_vs_nccl_1e-5_of_max"] = diff < 1e-5
    both = losses[0][0].clone()
    dist.all_reduce(both, op=dist.ReduceOp.MAX, group=group)
    checks["train_step_loss_same_on_every_rank"] = torch.equal(both, losses[0][0])

    # ---- evaluation sharded by user vs the whole tables on one GPU
    with torch.no_grad():
        u_glob, i_glob = e_glob[:N_USERS].contiguous(), e_glob[N_USERS:].contiguous()
        mask = graph.build_interaction_csr(u, i, N_USERS, N_ITEMS, device=dev)
        u0, u1 = part.users_of(rank)
        own_rows = own(e_glob)
        out_u, out_i = own_rows[:part.up], own_rows[part.up:]
        ptr_own = (mask.indptr[u0:u1 + 1] - mask.indptr[u0]).contiguous()
        idx_own = mask.indices[int(mask.indptr[u0]):int(mask.indptr[u1])].contiguous()
        users_local = torch.arange(0, u1 - u0, 3, device=dev, dtype=torch.int32)
        for mode in ("exact", "refquirk"):
            ids_s, sc_s = hdist.fullrank_topk_sharded(adj, out_u, out_i, users_local, ptr_own, idx_own, N_ITEMS, 20, mode=mode)
            ids_w, sc_w = evaluation.fullrank_topk(u_glob, i_glob, (users_local + u0).to(torch.int32), mask.indptr, mask.indices, 20, mode=mode)
            checks["eval_%s_sharded_eq_unsharded" % mode] = torch.equal(ids_s, ids_w) and torch.equal(sc_s, sc_w)
    torch.cuda.synchronize()
    flags = torch.tensor([int(bool(v)) for v in checks.values()], device=dev, dtype=torch.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)  # a check passes when it passes on every rank
    checks = {k: bool(f) for k, f in zip(checks.keys(), flags.tolist())}
    return {"ok": all(checks.values()), "world": world, "checks": checks,
            "graph": "%d users x %d items x %d interactions, emb %d" % (N_USERS, N_ITEMS, N_TRAIN, D)}


def scale_check(adj, d: int = 64, rounds: int = 3) -> dict:
    """The same question at the size of the job itself: on THIS rank's block of the benchmark graph, a propagation whose input
    table arrives through the fused exchange (peer / multicast stores issued by the producing kernel, device-side barrier) must
    give the bits of the same propagation fed by an NCCL ``all_gather``.  A store that has not landed when a peer starts reading
    shows up here as a differing row; every round uses fresh data and goes through a different slot of the pool."""
    dev = adj.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + adj.rank)
    ok_copy = ok_chain = True
    with torch.no_grad():
        for _ in range(rounds):
            x = torch.randn(adj.part.n_loc, d, device=dev, generator=gen)
            want1 = adj.k.spmm(adj.block, adj.all_gather(x))
            want2 = adj.k.spmm(adj.block, adj.all_gather(want1))
            full = adj.gathered(x)                                   # copy kernel -> every rank's table
            _, got1 = adj.spmm_published(full, None, want_local=True)  # propagation epilogue -> every rank's table
            got2 = adj.k.spmm(adj.block, adj.gathered(got1))
            ok_copy = ok_copy and torch.equal(got1, want1)
            ok_chain = ok_chain and torch.equal(got2, want2)
    flags = torch.tensor([int(ok_copy), int(ok_chain)], device=dev, dtype=torch.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=adj.group)
    return {"job_graph_copy_gather_eq_nccl": bool(flags[0]), "job_graph_fused_gather_eq_nccl": bool(flags[1]),
            "multicast": bool(getattr(adj, "multicast", False))}
