"""The body of the reference's training loop on the fused kernels.

``train_step`` is the per-batch body of ``LightGCN.train`` (model/graph/LightGCN.py:49-66), identical
in ``HGNN_HD3.train`` (model/graph/HGNN_HD3.py:138-160): full-graph propagation, gather of the batch
rows, BPR + L2/batch_size, backward, optimiser step.  The three ``.item()`` host synchronisations of the
reference are left to the caller: the losses come back as a device tensor.
"""
from __future__ import annotations

import torch

from .loss_torch import bpr_l2_from_tables, contrastLoss, unique_padded


def train_step(model, optimizer, user_idx, pos_idx, neg_idx, reg: float, batch_size: int, forward=None):
    """Returns the device tensor ``[rec_loss, reg_loss]`` (``reg_loss`` already divided by ``batch_size``)."""
    rec_user_emb, rec_item_emb = (forward or model)()[:2]
    rec_loss, reg_loss = bpr_l2_from_tables(rec_user_emb, rec_item_emb, user_idx, pos_idx, neg_idx, reg, batch_size)
    batch_loss = rec_loss + reg_loss
    optimizer.zero_grad(set_to_none=True)
    batch_loss.backward()
    optimizer.step()
    return torch.stack([rec_loss.detach(), reg_loss.detach()])


def train_step_hccf(model, optimizer, user_idx, pos_idx, neg_idx, temp: float, ss_rate: float, keep_rate: float, device_rng: bool = True,
                    static_shapes: bool = False):
    """Per-batch body of ``HCCF.train`` (model/graph/HCCF.py:79-95) with ``calcLosses`` (:59-68): edge-dropped GCN +
    learned-hyperedge propagation, BPR, and per layer ``contrastLoss`` between the detached GCN view and the hypergraph
    view for users and items.  SSL nodes = the batch's unique user / positive-item ids (see oracle/torch_path.py).
    ``device_rng``: draw the edge-drop mask on the GPU instead of the reference's CPU ``torch.rand(nnz)``.
    ``static_shapes``: ``loss_torch.unique_padded`` instead of ``torch.unique`` (no host sync, fixed shapes: capturable)."""
    n_users = model.data.n_users
    user_emb, item_emb, gcn_l, hyp_l = model(keep_rate=keep_rate, device_rng=device_rng)
    rec_loss, _ = bpr_l2_from_tables(user_emb, item_emb, user_idx, pos_idx, neg_idx, 0.0, 1.0)
    un, pn = (unique_padded(user_idx), unique_padded(pos_idx)) if static_shapes else (torch.unique(user_idx), torch.unique(pos_idx))
    ssl = 0
    for g, h in zip(gcn_l, hyp_l):
        g = g.detach()
        ssl = ssl + contrastLoss(g[:n_users], h[:n_users], un, temp) + contrastLoss(g[n_users:], h[n_users:], pn, temp)
    ssl = ssl * ss_rate
    optimizer.zero_grad(set_to_none=True)
    (rec_loss + ssl).backward()
    optimizer.step()
    return torch.stack([rec_loss.detach(), ssl.detach()])


class GraphedStep:
    """Any fixed-shape step ``fn(user_idx, pos_idx, neg_idx) -> losses`` captured into a CUDA graph (see GraphedTrainStep)."""

    def __init__(self, fn, batch: int, device, warmup: int = 3):
        self.u = torch.zeros(batch, dtype=torch.int64, device=device)
        self.p = torch.zeros(batch, dtype=torch.int64, device=device)
        self.n = torch.zeros(batch, dtype=torch.int64, device=device)
        self.batch = batch
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(self.u, self.p, self.n)
        torch.cuda.current_stream().wait_stream(side)
        from . import _lib

        before = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = fn(self.u, self.p, self.n)
        self.kernels_per_replay = _lib.launch_count() - before

    def __call__(self, user_idx, pos_idx, neg_idx):
        if user_idx.numel() != self.batch:
            raise ValueError("graphed step was captured for batch %d, got %d" % (self.batch, user_idx.numel()))
        self.u.copy_(user_idx, non_blocking=True)
        self.p.copy_(pos_idx, non_blocking=True)
        self.n.copy_(neg_idx, non_blocking=True)
        self.graph.replay()
        return self.losses


class GraphedTrainStep(GraphedStep):
    """``train_step`` captured once into a CUDA graph and replayed: for graphs whose propagation takes tens of microseconds
    (the L2-resident BASELINE shapes) the step is bound by ~25-60 kernel launches and the Python between them, not by the
    kernels.  The batch's index tensors are copied into static buffers, the whole forward / loss / backward / Adam sequence
    replays as one launch.  Requires a fixed batch size (the short last batch of an epoch runs through ``train_step``) and an
    optimizer built with ``capturable=True``."""

    def __init__(self, model, optimizer, reg: float, batch_size: int, batch: int, forward=None, warmup: int = 3):
        super().__init__(lambda u, p, n: train_step(model, optimizer, u, p, n, reg, batch_size, forward), batch,
                         next(model.parameters()).device, warmup)
