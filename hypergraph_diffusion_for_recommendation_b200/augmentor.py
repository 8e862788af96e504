"""SGL graph augmentation on the device -- ``data/augmentor.py`` + ``Interaction.convert_to_laplacian_mat``.

Reference (data/augmentor.py:11-42, data/ui_graph.py:86-93): ``node_dropout`` / ``edge_dropout`` sample with python's
``random.sample`` on scipy matrices, then ``convert_to_laplacian_mat`` rebuilds and re-normalises the ``(U+I)^2``
adjacency on the host -- twice per epoch in SGL.  Here the kept interactions are selected on the GPU (``torch.randperm`` /
``torch.rand``) and the normalised adjacency of the perturbed graph comes straight out of the device builder
(``graph.build_norm_adj``: radix sort + degrees + ``D^-1/2 A D^-1/2``), returned as the ``DeviceCSR`` the encoders take.
Same distribution as the reference (a uniform random subset of fixed size), different random stream.
"""
from __future__ import annotations

import torch

from . import graph


class GraphAugmentor(object):
    """Operates on the dense-id interaction list ``(user_idx, item_idx)`` (int tensors on the device)."""

    @staticmethod
    def edge_dropout(user_idx: torch.Tensor, item_idx: torch.Tensor, drop_rate: float, generator=None):
        """Keep ``int(E * (1 - drop_rate))`` interactions chosen uniformly without replacement (augmentor.py:32-42)."""
        e = int(user_idx.numel())
        keep = torch.randperm(e, device=user_idx.device, generator=generator)[:int(e * (1 - drop_rate))]
        return user_idx[keep], item_idx[keep]

    @staticmethod
    def node_dropout(user_idx: torch.Tensor, item_idx: torch.Tensor, n_users: int, n_items: int, drop_rate: float, generator=None):
        """Drop ``int(n * drop_rate)`` users and items chosen uniformly; interactions touching them vanish (augmentor.py:12-30)."""
        dev = user_idx.device
        ku = torch.ones(n_users, dtype=torch.bool, device=dev)
        ki = torch.ones(n_items, dtype=torch.bool, device=dev)
        ku[torch.randperm(n_users, device=dev, generator=generator)[:int(n_users * drop_rate)]] = False
        ki[torch.randperm(n_items, device=dev, generator=generator)[:int(n_items * drop_rate)]] = False
        m = ku[user_idx.long()] & ki[item_idx.long()]
        return user_idx[m], item_idx[m]


def convert_to_laplacian_mat(user_idx: torch.Tensor, item_idx: torch.Tensor, n_users: int, n_items: int) -> graph.DeviceCSR:
    """``Interaction.convert_to_laplacian_mat`` (data/ui_graph.py:86-93) of the perturbed interaction list."""
    return graph.build_norm_adj(user_idx, item_idx, n_users, n_items, device=user_idx.device)


def random_graph_augment(user_idx, item_idx, n_users, n_items, aug_type: int, drop_rate: float, generator=None) -> graph.DeviceCSR:
    """``SGL_Encoder.random_graph_augment`` (model/graph/SGL.py:138-145): aug_type 0 = node dropout, 1 / 2 = edge dropout."""
    if aug_type == 0:
        u, i = GraphAugmentor.node_dropout(user_idx, item_idx, n_users, n_items, drop_rate, generator)
    else:
        u, i = GraphAugmentor.edge_dropout(user_idx, item_idx, drop_rate, generator)
    return convert_to_laplacian_mat(u, i, n_users, n_items)
