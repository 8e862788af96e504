"""world_size-2 gloo tests (CPU) of the row-partitioned path: partition maps, the rank-local adjacency block,
the gather / autograd wiring of the sharded propagation and one sharded training step.  The CUDA kernels are
replaced by a CPU kernel set built on the oracle's sequential-FMA matrix product (tests may use the oracle);
everything else is the product code of hypergraph_diffusion_for_recommendation_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from hypergraph_diffusion_for_recommendation_b200 import dist as hdist
from oracle import hgr_oracle as O
from oracle import torch_path as T

N_USERS, N_ITEMS, D = 37, 53, 64


def graph():
    from hypergraph_diffusion_for_recommendation_b200.synth import powerlaw_interactions

    g = powerlaw_interactions(N_USERS, N_ITEMS, 600, seed=11)
    return g.train_u, g.train_i


class CpuKernels:
    """Same three entry points as dist.LibhgrKernels, on the CPU."""

    @staticmethod
    def make_block(indptr, indices, values, shape):
        return (indptr.numpy().astype(np.int64), indices.numpy().astype(np.int64), values.numpy(), shape)

    @staticmethod
    def spmm(block, x, ep=None):
        indptr, indices, values, _ = block
        y = torch.from_numpy(O.spmm(indptr, indices, values, x.detach().numpy()))
        ep = ep or {}
        if ep.get("pre") is not None:
            ep["pre"].copy_(y)
        if ep.get("slope") is not None:
            y = F.leaky_relu(y, ep["slope"])
        if ep.get("gamma") is not None:
            y = F.layer_norm(y, (y.shape[1],), ep["gamma"], ep["beta"], ep.get("eps", 1e-5))
        if ep.get("residual") is not None:
            y = y + ep["residual"]
        if ep.get("addends"):
            y = (y + sum(ep["addends"])) * ep.get("scale", 1.0)
        return y.detach()

    @staticmethod
    def leaky_ln_bwd(pre, dy, gamma, eps, slope):
        with torch.enable_grad():  # called from inside an autograd backward
            p = pre.detach().clone().requires_grad_(True)
            g = gamma.detach().clone().requires_grad_(True) if gamma is not None else None
            b = torch.zeros_like(g, requires_grad=True) if g is not None else None
            y = F.leaky_relu(p, slope) if slope is not None else p
            if g is not None:
                y = F.layer_norm(y, (y.shape[1],), g, b, eps)
            grads = torch.autograd.grad(y, [p] + ([g, b] if g is not None else []), dy)
        return grads[0], (grads[1] if g is not None else None), (grads[2] if g is not None else None)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(fn, world=2):
    port = free_port()
    mp.spawn(_entry, args=(world, port, fn), nprocs=world, join=True)


def _entry(rank, world, port, fn):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


# ------------------------------------------------------------------------------------------------ partition (no comm)
def test_partition_maps_cover_every_node_once():
    part = hdist.Partition(N_USERS, N_ITEMS, 4)
    pu = part.perm_user(torch.arange(N_USERS))
    pi = part.perm_item(torch.arange(N_ITEMS))
    allp = torch.cat([pu, pi])
    assert allp.unique().numel() == N_USERS + N_ITEMS and int(allp.max()) < part.n_glob
    for r in range(4):
        u0, u1 = part.users_of(r)
        assert ((pu[u0:u1] >= r * part.n_loc) & (pu[u0:u1] < r * part.n_loc + part.up)).all()
        i0, i1 = part.items_of(r)
        assert ((pi[i0:i1] >= r * part.n_loc + part.up) & (pi[i0:i1] < (r + 1) * part.n_loc)).all()


def test_balanced_partition_cuts_by_nonzeros():
    """Partition.balanced: contiguous ranges, every node owned once, rank blocks with (nearly) equal nonzeros even when a few
    items hold most interactions -- where equal-count ranges are off by tens of per cent."""
    rng = np.random.default_rng(9)
    n_u, n_i, world = 4000, 900, 8
    deg_i = (rng.pareto(1.1, n_i) * 20 + 1).astype(np.int64)
    deg_i[17] = deg_i.sum() // 12  # one item with ~8 % of everything
    deg_u = rng.multinomial(int(deg_i.sum()), np.full(n_u, 1.0 / n_u)).astype(np.int64)
    part = hdist.Partition.balanced(n_u, n_i, world, torch.from_numpy(deg_u), torch.from_numpy(deg_i))
    flat = hdist.Partition(n_u, n_i, world)
    assert not part.uniform and flat.uniform
    pu, pi = part.perm_user(torch.arange(n_u)), part.perm_item(torch.arange(n_i))
    allp = torch.cat([pu, pi])
    assert allp.unique().numel() == n_u + n_i and int(allp.max()) < part.n_glob

    def load(p):
        return np.array([deg_u[slice(*p.users_of(r))].sum() + deg_i[slice(*p.items_of(r))].sum() for r in range(world)], dtype=np.float64)

    for r in range(world):
        u0, u1 = part.users_of(r)
        i0, i1 = part.items_of(r)
        assert u1 - u0 <= part.up and i1 - i0 <= part.ip
        assert ((pu[u0:u1] >= r * part.n_loc) & (pu[u0:u1] < r * part.n_loc + part.up)).all()
        assert ((pi[i0:i1] >= r * part.n_loc + part.up) & (pi[i0:i1] < (r + 1) * part.n_loc)).all()
        assert (torch.diff(pu[u0:u1]) == 1).all() and (torch.diff(pi[i0:i1]) == 1).all()  # monotonic inside a range
    lb, lf = load(part), load(flat)
    assert lb.sum() == lf.sum() == deg_u.sum() + deg_i.sum()
    assert lb.max() / lb.mean() < 1.0 + 0.6 * (lf.max() / lf.mean() - 1.0)  # the heavy item cannot be split, the rest evens out
    with pytest.raises(ValueError):
        hdist.Partition(n_u, n_i, world, user_bounds=(0, 5, 3) + (n_u,) * 6)


@pytest.mark.parametrize("world,balanced", [(1, False), (2, False), (3, False), (8, False), (3, True), (8, True)])
def test_local_blocks_reassemble_the_global_adjacency(world, balanced):
    u, i = graph()
    indptr, indices, values = O.build_norm_adj(u, i, N_USERS, N_ITEMS)
    if balanced:
        part = hdist.Partition.balanced(N_USERS, N_ITEMS, world, torch.bincount(torch.from_numpy(u).long(), minlength=N_USERS),
                                        torch.bincount(torch.from_numpy(i).long(), minlength=N_ITEMS))
    else:
        part = hdist.Partition(N_USERS, N_ITEMS, world)
    perm = torch.cat([part.perm_user(torch.arange(N_USERS)), part.perm_item(torch.arange(N_ITEMS))]).numpy()
    seen = 0
    for r in range(world):
        lp, li, lv = hdist.local_block(part, r, torch.from_numpy(u), torch.from_numpy(i))
        lp, li, lv = lp.numpy(), li.numpy(), lv.numpy()
        u0, u1 = part.users_of(r)
        i0, i1 = part.items_of(r)
        rows = list(range(u0, u1)) + [None] * (part.up - (u1 - u0)) + [N_USERS + k for k in range(i0, i1)]
        rows += [None] * (part.n_loc - len(rows))
        for loc, g in enumerate(rows):
            a, b = lp[loc], lp[loc + 1]
            if g is None:
                assert a == b
                continue
            ga, gb = indptr[g], indptr[g + 1]
            # same nonzeros in the same (original ascending) order, columns renumbered by the permutation
            assert np.array_equal(li[a:b], perm[indices[ga:gb]])
            assert np.array_equal(lv[a:b].view(np.uint32), values[ga:gb].view(np.uint32))
            seen += b - a
    assert seen == indices.size


# ------------------------------------------------------------------------------------------------ world_size 2
def _setup(rank, world):
    u, i = graph()
    ctx = hdist.build_partitioned(torch.from_numpy(u), torch.from_numpy(i), N_USERS, N_ITEMS, rank, world, kernels=CpuKernels)
    csr = O.build_norm_adj(u, i, N_USERS, N_ITEMS)
    part = ctx.part
    rng = np.random.default_rng(5)
    e_glob = torch.from_numpy((rng.standard_normal((N_USERS + N_ITEMS, D)) * 0.1).astype(np.float32))
    g_glob = torch.from_numpy(rng.standard_normal((N_USERS + N_ITEMS, D)).astype(np.float32))
    perm = torch.cat([part.perm_user(torch.arange(N_USERS)), part.perm_item(torch.arange(N_ITEMS))])

    def own(t):  # global [N, D] -> this rank's [n_loc, D] block
        full = torch.zeros(part.n_glob, t.shape[1])
        full[perm] = t
        return full[rank * part.n_loc:(rank + 1) * part.n_loc].clone()

    def mask_own():
        m = torch.zeros(part.n_glob, dtype=torch.bool)
        m[perm] = True
        return m[rank * part.n_loc:(rank + 1) * part.n_loc]

    return ctx, csr, e_glob, g_glob, own, mask_own()


def _w_lightgcn(rank, world):
    ctx, csr, e_glob, g_glob, own, live = _setup(rank, world)
    x = own(e_glob).requires_grad_(True)
    out = ctx.adj.lightgcn_propagate(x, 3)
    ue, ie = O.lgcn_forward(csr, e_glob[:N_USERS].numpy(), e_glob[N_USERS:].numpy(), 3)
    want = own(torch.from_numpy(np.concatenate([ue, ie])))
    # (the layer-mean readout adds the layers in a different order than torch.mean: tolerance, not bits)
    assert rel(out.detach()[live], want[live]) < 1e-6
    (out * own(g_glob)).sum().backward()
    adj = T.coo_from_csr(*csr, (N_USERS + N_ITEMS,) * 2)
    eg = e_glob.clone().requires_grad_(True)
    acc, cur = eg, eg
    for _ in range(3):
        cur = torch.sparse.mm(adj, cur)
        acc = acc + cur
    ((acc / 4) * g_glob).sum().backward()
    assert rel(x.grad[live], own(eg.grad)[live]) < 1e-5
    # plain propagation: column order is preserved, so the sharded rows carry the same BITS as the unsharded oracle
    y = ctx.adj.spmm(own(e_glob))
    assert torch.equal(y[live], own(torch.from_numpy(O.spmm(*csr, e_glob.numpy())))[live])


def _w_hgconv(rank, world):
    ctx, csr, e_glob, g_glob, own, live = _setup(rank, world)
    torch.manual_seed(3)
    gamma0, beta0 = torch.randn(D), torch.randn(D)
    x = own(e_glob).requires_grad_(True)
    gamma, beta = gamma0.clone().requires_grad_(True), beta0.clone().requires_grad_(True)
    y = ctx.adj.hgconv(x, 0.5, gamma, beta, residual=x)
    (y * own(g_glob))[live].sum().backward()
    hdist.sync_replicated_grads(torch.nn.ParameterDict({"g": torch.nn.Parameter(gamma0)}), group=None)  # no grads: no-op
    for p in (gamma, beta):
        dist.all_reduce(p.grad)
    adj = T.coo_from_csr(*csr, (N_USERS + N_ITEMS,) * 2)
    eg = e_glob.clone().requires_grad_(True)
    gg, bg = gamma0.clone().requires_grad_(True), beta0.clone().requires_grad_(True)
    yg = F.layer_norm(F.leaky_relu(torch.sparse.mm(adj, torch.sparse.mm(adj, eg)), 0.5), (D,), gg, bg) + eg
    (yg * g_glob).sum().backward()
    assert rel(y.detach()[live], own(yg.detach())[live]) < 1e-5
    assert rel(x.grad[live], own(eg.grad)[live]) < 1e-5
    assert rel(gamma.grad, gg.grad) < 1e-5 and rel(beta.grad, bg.grad) < 1e-5


def _w_train_step(rank, world):
    from hypergraph_diffusion_for_recommendation_b200 import encoders

    ctx, csr, e_glob, g_glob, own, live = _setup(rank, world)
    part = ctx.part
    torch.manual_seed(0)
    model = encoders.LGCN_Encoder.__new__(encoders.LGCN_Encoder)
    torch.nn.Module.__init__(model)
    model.data, model.layers, model.sparse_norm_adj = ctx.data, 2, ctx.adj
    blk = own(e_glob)
    model.embedding_dict = torch.nn.ParameterDict({"user_emb": torch.nn.Parameter(blk[:part.up].clone()),
                                                   "item_emb": torch.nn.Parameter(blk[part.up:].clone())})
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    rng = np.random.default_rng(9)
    tu, ti = graph()
    pick = rng.integers(0, tu.size, 128)
    bu, bp, bn = torch.from_numpy(tu[pick]), torch.from_numpy(ti[pick]), torch.from_numpy(rng.integers(0, N_ITEMS, 128))

    def loss_fn(ut, it, u, p, n, reg, bs):
        return T.bpr_loss(ut[u], it[p], it[n]), T.l2_reg_loss(reg, ut[u], it[p], it[n]) / bs

    losses = hdist.train_step(model, opt, ctx.adj, bu, bp, bn, 0.01, 128, loss_fn=loss_fn)
    # single-process reference: the restated reference model (oracle/torch_path.py) on the global graph
    adj = T.coo_from_csr(*csr, (N_USERS + N_ITEMS,) * 2)
    ref = T.LGCN(adj, N_USERS, N_ITEMS, D, 2)
    ref.load_state_dict({"embedding_dict.user_emb": e_glob[:N_USERS], "embedding_dict.item_emb": e_glob[N_USERS:]})
    ropt = torch.optim.Adam(ref.parameters(), lr=0.01)
    rl = T.train_step(ref, ropt, bu, bp, bn, 0.01, 128)
    assert abs(float(losses[0]) - rl[1]) < 1e-5 * abs(rl[1]) and abs(float(losses[1]) - rl[2]) < 1e-5 * abs(rl[2])
    want = own(torch.cat([ref.embedding_dict["user_emb"].detach(), ref.embedding_dict["item_emb"].detach()]))
    got = torch.cat([model.embedding_dict["user_emb"].detach(), model.embedding_dict["item_emb"].detach()])
    assert rel(got[live], want[live]) < 1e-5


def test_sharded_lightgcn_forward_bit_exact_and_backward():
    run_world(_w_lightgcn)


def test_sharded_hgconv_with_layernorm_and_residual():
    run_world(_w_hgconv)


def test_sharded_training_step_matches_single_process():
    run_world(_w_train_step)
