"""Synthetic power-law interaction graphs of the BASELINE.json shapes.

The reference ships no dataset (``.gitignore:4-5`` of the reference ignores ``dataset/``), so every
parity test and benchmark runs on graphs drawn here:

* user activity  ~ rank^-0.8, item popularity ~ rank^-1.0, each under a random permutation of ids,
* every user and every item receives at least one interaction,
* pairs are unique and topped up to exactly ``n_total`` interactions,
* a random 75 / 25 train / test split mirroring the reference's ``dataset_util.py:20-37``.

``powerlaw_interactions`` is the numpy generator used for C1-C4 (it also writes the reference's
text format, ``data/loader.py:24-38``: one header line, then ``user<TAB>item``);
``powerlaw_interactions_device`` draws the same distribution with torch ops on the GPU for the
10 M x 2 M x 1 B shape where a host generator would take minutes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# (n_users, n_items, n_train_interactions) per BASELINE.json config; SURVEY.md section 8 table.
SHAPES = {
    "c1_lastfm": (1891, 14777, 70_000),
    "c2_ml1m": (6040, 3706, 750_000),
    "c3_gowalla": (30_000, 41_000, 1_000_000),
    "c4_amazon_book": (52_000, 92_000, 3_000_000),
    "c5_1b": (10_000_000, 2_000_000, 1_000_000_000),
}


@dataclass
class SynthGraph:
    n_users: int
    n_items: int
    train_u: np.ndarray  # int64 [E_train] dense user ids
    train_i: np.ndarray  # int64 [E_train] dense item ids
    test_u: np.ndarray
    test_i: np.ndarray


def _zipf_cdf(n: int, alpha: float, rng: np.random.Generator) -> tuple[np.ndarray, np.ndarray]:
    w = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    perm = rng.permutation(n)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return cdf, perm


def powerlaw_interactions(n_users: int, n_items: int, n_train: int, seed: int = 1234,
                          test_ratio: float = 0.25) -> SynthGraph:
    """Draw ``n_train / (1 - test_ratio)`` unique (user, item) pairs and split them."""
    rng = np.random.default_rng(seed)
    n_total = int(round(n_train / (1.0 - test_ratio)))
    if n_total > n_users * n_items:
        raise ValueError("more interactions requested than user x item pairs")
    ucdf, uperm = _zipf_cdf(n_users, 0.8, rng)
    icdf, iperm = _zipf_cdf(n_items, 1.0, rng)

    # coverage: one interaction for every user and one for every item
    cu = np.arange(n_users, dtype=np.int64)
    ci = iperm[np.minimum(np.searchsorted(icdf, rng.random(n_users)), n_items - 1)]
    di = np.arange(n_items, dtype=np.int64)
    du = uperm[np.minimum(np.searchsorted(ucdf, rng.random(n_items)), n_users - 1)]
    keys = np.unique(np.concatenate([cu * n_items + ci, du * n_items + di]))
    while keys.size < n_total:
        need = n_total - keys.size
        m = int(need * 1.3) + 1024
        u = uperm[np.minimum(np.searchsorted(ucdf, rng.random(m)), n_users - 1)]
        i = iperm[np.minimum(np.searchsorted(icdf, rng.random(m)), n_items - 1)]
        new = np.setdiff1d(np.unique(u.astype(np.int64) * n_items + i), keys, assume_unique=True)
        if new.size > need:
            new = rng.permutation(new)[:need]
        keys = np.union1d(keys, new)
    if keys.size > n_total:  # coverage pairs overshoot only on tiny shapes
        keys = rng.permutation(keys)[:n_total]
    keys = rng.permutation(keys)
    n_test = n_total - n_train
    test, train = keys[:n_test], keys[n_test:]
    return SynthGraph(n_users, n_items, train // n_items, train % n_items, test // n_items, test % n_items)


def write_reference_files(g: SynthGraph, root: str, dataset: str = "lastfm") -> str:
    """Write ``train.txt`` / ``test.txt`` (+ the stub ``lastfm.kg`` that the reference's
    ``SELFRec.py:18`` always opens) under ``root/dataset/<dataset>/``; raw item ids are offset by
    ``n_users`` so the two id spaces do not collide."""
    d = os.path.join(root, "dataset", dataset)
    os.makedirs(d, exist_ok=True)
    for name, (u, i) in {"train.txt": (g.train_u, g.train_i), "test.txt": (g.test_u, g.test_i)}.items():
        with open(os.path.join(d, name), "w") as f:
            f.write("user\titem\n")
            np.savetxt(f, np.stack([u, i + g.n_users], 1), fmt="%d", delimiter="\t")
    kg = os.path.join(root, "dataset", "lastfm")
    os.makedirs(kg, exist_ok=True)
    with open(os.path.join(kg, "lastfm.kg"), "w") as f:
        f.write("head\trelation\ttail\n0\t0\t1\n")
    return d


def reference_dense_ids(train_u: np.ndarray, train_i: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Dense ids in order of first appearance, the way ``Interaction.__generate_set``
    (reference ``data/ui_graph.py:43-68``) numbers users and items. Returns
    ``(dense_u, dense_i, id2user, id2item)``."""
    def first_seen(x):
        uniq, first, inv = np.unique(x, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        return rank[inv], uniq[order]
    du, id2user = first_seen(train_u)
    di, id2item = first_seen(train_i)
    return du, di, id2user, id2item


def powerlaw_interactions_device(n_users: int, n_items: int, n_train: int, device, seed: int = 1234,
                                 chunk: int = 1 << 27, cover: bool = True, perm_seed: int | None = None):
    """Device generator for shapes where the numpy path is too slow (C5). Returns int32 tensors
    ``(train_u, train_i)`` of unique pairs; the test split is drawn by the caller from the same
    distribution. Sampling is inverse-CDF through ``torch.searchsorted``."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    uw = torch.arange(1, n_users + 1, device=device, dtype=torch.float64).pow_(-0.8)
    iw = torch.arange(1, n_items + 1, device=device, dtype=torch.float64).pow_(-1.0)
    ucdf = torch.cumsum(uw, 0)
    ucdf /= ucdf[-1].clone()
    icdf = torch.cumsum(iw, 0)
    icdf /= icdf[-1].clone()
    del uw, iw
    # popularity ranks -> ids.  ``perm_seed`` = the seed of another call: same popular users / items, fresh draws
    # (held-out interactions of a graph generated with that seed)
    pgen = gen
    if perm_seed is not None:
        pgen = torch.Generator(device=device)
        pgen.manual_seed(perm_seed)
    uperm = torch.randperm(n_users, device=device, generator=pgen)
    iperm = torch.randperm(n_items, device=device, generator=pgen)

    def draw(m):
        ru = torch.rand(m, device=device, dtype=torch.float64, generator=gen)
        u = uperm[torch.searchsorted(ucdf, ru).clamp_(max=n_users - 1)]
        del ru
        ri = torch.rand(m, device=device, dtype=torch.float64, generator=gen)
        i = iperm[torch.searchsorted(icdf, ri).clamp_(max=n_items - 1)]
        del ri
        return u * n_items + i

    if cover:  # one interaction for every user and for every item (training graphs)
        cu = torch.arange(n_users, device=device, dtype=torch.int64)
        ri = torch.rand(n_users, device=device, dtype=torch.float64, generator=gen)
        cov1 = cu * n_items + iperm[torch.searchsorted(icdf, ri).clamp_(max=n_items - 1)]
        di = torch.arange(n_items, device=device, dtype=torch.int64)
        ru = torch.rand(n_items, device=device, dtype=torch.float64, generator=gen)
        cov2 = uperm[torch.searchsorted(ucdf, ru).clamp_(max=n_users - 1)] * n_items + di
        keys = torch.unique(torch.cat([cov1, cov2]))
        del cu, ri, di, ru, cov1, cov2
    else:
        keys = torch.empty(0, dtype=torch.int64, device=device)
    while keys.numel() < n_train:
        need = n_train - keys.numel()
        m = min(int(need * 1.5) + 4096, chunk * 8)
        parts = [draw(min(chunk, m - s)) for s in range(0, m, chunk)]
        keys = torch.unique(torch.cat([keys] + parts))
        del parts
    if keys.numel() > n_train:
        # drop a random subset of the surplus (keeps the coverage pairs with overwhelming probability)
        keep = torch.randperm(keys.numel(), device=device, generator=gen)[:n_train]
        keys = keys[keep.sort().values]
    u = torch.div(keys, n_items, rounding_mode="floor").to(torch.int32)
    i = (keys % n_items).to(torch.int32)
    return u, i


def norm_adj_from_pairs_torch(u, i, n_users: int, n_items: int, chunk_nnz=None):
    """D^-1/2 A D^-1/2 of the bipartite graph of UNIQUE device pairs ``(u, i)`` as a ``DeviceCSR``,
    assembled with torch ops (sort + bincount).  Benchmark plumbing for graphs too large for the host
    generator; the product builder is ``graph.build_norm_adj`` (csrc/graph_build.cu), which also sums
    duplicates and takes the host's ``np.power`` lookup table for bit-exact values."""
    import torch

    from .graph import DeviceCSR

    n = n_users + n_items
    u = u.to(torch.int64)
    i = i.to(torch.int64)
    deg_u = torch.bincount(u, minlength=n_users)
    deg_i = torch.bincount(i, minlength=n_items)
    deg = torch.cat([deg_u, deg_i])
    d = deg.to(torch.float32).rsqrt()
    d[deg == 0] = 0
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=u.device)
    torch.cumsum(deg, 0, out=indptr[1:])
    # user rows: items ascending; item rows: users ascending
    key = u * n_items + i
    key, _ = torch.sort(key)
    ru, ci = torch.div(key, n_items, rounding_mode="floor"), key % n_items
    del key
    key2, _ = torch.sort(i * n_users + u)
    ri, cu = torch.div(key2, n_users, rounding_mode="floor"), key2 % n_users
    del key2
    indices = torch.cat([(ci + n_users).to(torch.int32), cu.to(torch.int32)])
    values = torch.cat([d[ru] * d[ci + n_users], d[ri + n_users] * d[cu]])
    return DeviceCSR(indptr, indices, values, (n, n), symmetric=True, chunk_nnz=chunk_nnz)
