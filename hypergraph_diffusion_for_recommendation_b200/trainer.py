"""The body of the reference's training loop on the fused kernels.

``train_step`` is the per-batch body of ``LightGCN.train`` (model/graph/LightGCN.py:49-66), identical
in ``HGNN_HD3.train`` (model/graph/HGNN_HD3.py:138-160): full-graph propagation, gather of the batch
rows, BPR + L2/batch_size, backward, optimiser step.  The three ``.item()`` host synchronisations of the
reference are left to the caller: the losses come back as a device tensor.
"""
from __future__ import annotations

import torch

from .loss_torch import bpr_l2_from_tables


def train_step(model, optimizer, user_idx, pos_idx, neg_idx, reg: float, batch_size: int, forward=None):
    """Returns the device tensor ``[rec_loss, reg_loss]`` (``reg_loss`` already divided by ``batch_size``)."""
    rec_user_emb, rec_item_emb = (forward or model)()[:2]
    rec_loss, reg_loss = bpr_l2_from_tables(rec_user_emb, rec_item_emb, user_idx, pos_idx, neg_idx, reg, batch_size)
    batch_loss = rec_loss + reg_loss
    optimizer.zero_grad(set_to_none=True)
    batch_loss.backward()
    optimizer.step()
    return torch.stack([rec_loss.detach(), reg_loss.detach()])
