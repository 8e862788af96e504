"""BPR pairwise batches on the device -- the drop-in for ``util.sampler.next_batch_pairwise``.

Reference (util/sampler.py:237-264): ``shuffle(training_data)``; per batch, for every positive ``(user, item)`` draw
``random.choice(item_list)`` until the item is not in ``training_set_u[user]``; yield CPU ``LongTensor``s
``(u_idx, i_idx, j_idx)`` of dense ids.  Here the shuffle is one ``torch.randperm`` on the device per epoch and the
negatives come from ``hgr_bpr_sample`` (csrc/sampler.cu: Philox + rejection against the user's training row); the
generator yields device LongTensors with the same meaning and order.  The random streams differ from python's, so
runs that must reproduce the reference's triples replay them instead (every loss entry point takes any index tensors).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class PairwiseSampler:
    """Training pairs + the training matrix (user -> sorted items) resident in HBM."""

    def __init__(self, train_u, train_i, n_users: int, n_items: int, device="cuda", seed: int = 0):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.HgrError("the sampler runs on the GPU (no CPU path)")
        self.edge_u = torch.as_tensor(train_u).to(device=dev, dtype=torch.int32).contiguous()
        self.edge_i = torch.as_tensor(train_i).to(device=dev, dtype=torch.int32).contiguous()
        self.n_users, self.n_items, self.device = int(n_users), int(n_items), dev
        key = torch.unique(self.edge_u.to(torch.int64) * n_items + self.edge_i.to(torch.int64))  # sorted, duplicates dropped
        self.train_indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=dev)
        torch.cumsum(torch.bincount(torch.div(key, n_items, rounding_mode="floor"), minlength=n_users), 0, out=self.train_indptr[1:])
        self.train_indices = (key % n_items).to(torch.int32)
        self.seed, self.drawn = int(seed), 0
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(int(seed))
        self.gave_up = torch.zeros(1, dtype=torch.int32, device=dev)

    @classmethod
    def from_interaction(cls, data, device="cuda", seed: int = 0):
        """From the reference's ``Interaction`` (data/ui_graph.py): ``training_data`` rows are ``[raw_user, raw_item, w]``."""
        if hasattr(data, "dense_training_pairs"):  # the array-backed façade (data.Interaction)
            u, i = data.dense_training_pairs()
            return cls(u, i, data.n_users, data.n_items, device=device, seed=seed)
        u = np.fromiter((data.user[int(e[0])] for e in data.training_data), dtype=np.int64, count=len(data.training_data))
        i = np.fromiter((data.item[int(e[1])] for e in data.training_data), dtype=np.int64, count=len(data.training_data))
        return cls(u, i, data.n_users, data.n_items, device=device, seed=seed)

    def batch(self, perm, offset: int, size: int, n_negs: int = 1):
        dev = self.device
        u = torch.empty(size, dtype=torch.int64, device=dev)
        p = torch.empty(size, dtype=torch.int64, device=dev)
        n = torch.empty(size * n_negs, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().hgr_bpr_sample(self.edge_u.data_ptr(), self.edge_i.data_ptr(), self.edge_u.numel(), _lib.ptr(perm), offset,
                                             size, n_negs, self.train_indptr.data_ptr(), self.train_indices.data_ptr(), self.n_items,
                                             self.seed, self.drawn, u.data_ptr(), p.data_ptr(), n.data_ptr(), self.gave_up.data_ptr(),
                                             _lib.stream_ptr()))
        self.drawn += size * n_negs
        return u, p, n

    def epoch(self, batch_size: int, n_negs: int = 1):
        """One pass over the shuffled training pairs; the last batch is short, as in the reference."""
        n = int(self.edge_u.numel())
        perm = torch.randperm(n, device=self.device, generator=self.gen)
        for off in range(0, n, batch_size):
            yield self.batch(perm, off, min(batch_size, n - off), n_negs)


def next_batch_pairwise(data, batch_size, n_negs=1, device=None):
    """Reference signature (util/sampler.py:237).  ``data`` is an ``Interaction`` or a ``PairwiseSampler``; the sampler is
    cached on the data object so the training matrix is uploaded once, not rebuilt every batch."""
    s = data if isinstance(data, PairwiseSampler) else getattr(data, "_hgr_sampler", None)
    if s is None:
        s = PairwiseSampler.from_interaction(data, device=device or "cuda")
        data._hgr_sampler = s
    yield from s.epoch(batch_size, n_negs)
