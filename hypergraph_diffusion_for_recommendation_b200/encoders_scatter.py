"""Encoders built on the SCATTER form of the hypergraph message passing.

===============================  =======================================================================================
``EquivSetGNN2``                  model/layers/layers2/EquivSetGNN2.py:32-103 (+ ``generate_V_E`` :105-133): ``lin_in`` -> scatter
                                  convolution (``EquivSetConv2.py:85-100``) -> ``W``.  The reference rebuilds ``(V, E)`` with
                                  ``torch.nonzero(dense (U+I)^2 matrix > 0)`` in every forward; here the incidence pair is a
                                  ``graph.Incidence`` built once from the sparse adjacency (``graph.incidence_from_csr``).
``LocalAwareEncoderHD4``          ``LocalAwareEncoder`` of model/graph/HGNN_HD4.py:337-405 (state_dict keys unchanged)
``HCCFDiffusionEncoder``          ``HCCFEncoder`` of model/graph/HCCF_diffusion.py:131-217: the hypergraph branch scatters over the
                                  SIGN pattern of the learned incidence ``E W`` (:382-402), which is dense (about half of its
                                  ``n x hyper_dim`` entries), so it runs as two tall-and-skinny products with a 0/1 matrix
                                  (``ops.pattern_mean_conv``) instead of ``nonzero`` + ``torch_scatter`` over n K / 2 pairs
``EquivSetConvAttention``         ``EquivSetConv`` of model/graph/HD2.py:589-643: one attention weight per (vertex, hyperedge) pair
===============================  =======================================================================================

Same constructor arguments, return values and ``state_dict`` keys as the reference classes, so checkpoints and the training
loops of ``HGNN_HD4`` / ``HCCF_diffusion`` carry over; the class shells come from ``encoders`` (subclassed, not restated).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import graph, ops
from .encoders import EquivSetConvScatter, HCCFEncoder, HGCNConv, LocalAwareEncoder, _adjacency_of


class EquivSetGNN2(nn.Module):
    def __init__(self, num_features, args, dense_hypergraph=None, data=None):
        super().__init__()
        hid = args['MLP_hidden']
        pick = lambda k: args['MLP_num_layers'] if args[k] < 0 else args[k]
        self.act = {'Id': nn.Identity(), 'relu': nn.ReLU(), 'prelu': nn.PReLU()}[args['activation']]
        self.p = float(args['dropout'])
        self.nlayer = args['All_num_layers']
        self.lin_in = nn.Linear(num_features, hid)
        self.conv = EquivSetConvScatter(hid, hid, mlp1_layers=args['MLP_num_layers'], mlp2_layers=pick('MLP2_num_layers'),
                                        mlp3_layers=pick('MLP3_num_layers'), alpha=args['restart_alpha'], aggr=args['aggregate'],
                                        dropout=args['dropout'], normalization=args['normalization'], input_norm=args['AllSet_input_norm'],
                                        data=data)

    def reset_parameters(self):
        self.lin_in.reset_parameters()
        self.conv.reset_parameters()

    def message(self, x, where, x0):
        return self.conv.forward_inc(x, where, x0)

    def forward(self, x, where, n_nodes=None):
        """``where``: the ``graph.Incidence`` to scatter over (the reference passes the dense matrix and derives (V, E) from it)."""
        drop = lambda t: F.dropout(t, self.p, self.training)
        x = ops.linear(drop(x), self.lin_in.weight, self.lin_in.bias, relu=True)
        x0 = x
        for _ in range(self.nlayer):
            x = self.act(self.message(drop(x), where, x0))
        return drop(x)


class LocalAwareEncoderHD4(LocalAwareEncoder):
    """Layers ``k < L - 1``: ``EquivSetGNN2`` over the star expansion of the bipartite adjacency ``+ res``; last layer:
    ``lns[0](HGCNConv(adj, act=False)) + res`` in one fused two-stage propagation (HGNN_HD4.py:391-405)."""

    def __init__(self, data, emb_size, hyper_size, n_layers, leaky, drop_rate, device=None, use_self_att=False):
        nn.Module.__init__(self)
        self.data, self.latent_size, self.hyper_size, self.layers = data, emb_size, hyper_size, n_layers
        self.sparse_norm_adj = _adjacency_of(data)
        self.edhnn_args = self.init_edhnn_config(hyper_size)
        self.hgcn_layers = nn.ModuleList([HGCNConv(leaky=0.5) for _ in range(n_layers)])
        self.edhnn_layers = nn.ModuleList([EquivSetGNN2(hyper_size, self.edhnn_args, None, data) for _ in range(n_layers)])
        self.lns = nn.ModuleList([nn.LayerNorm(hyper_size) for _ in range(n_layers)])
        self.edhnn_ui_n = data.n_users + data.n_items
        self._incidence = None

    @property
    def incidence(self):
        # ui_adj and norm_adj share their sparsity pattern and every stored value is positive: nonzero(ui_adj > 0) is the pattern
        if self._incidence is None:
            self._incidence = graph.incidence_from_csr(self.sparse_norm_adj)
        return self._incidence

    def forward(self, ego_embeddings, sparse_norm_adj):
        res = ego_embeddings
        for k in range(self.layers):
            if k != self.layers - 1:
                ego_embeddings = self.edhnn_layers[k](ego_embeddings, self.incidence, self.edhnn_ui_n) + res
            else:
                ego_embeddings = self.hgcn_layers[0](sparse_norm_adj, ego_embeddings, act=False, ln=self.lns[0], residual=res)
        return ego_embeddings[:self.data.n_users], ego_embeddings[self.data.n_users:]


class _EquivSetGNNPattern(EquivSetGNN2):
    """``EquivSetGNN`` of HCCF_diffusion (:310-402): ``where`` is the dense 0/1 incidence ``[n, hyper_dim]``."""

    def message(self, x, where, x0):
        conv = self.conv
        xv = ops.pattern_mean_conv(where, conv.W1(x))
        return conv.W(xv if conv.alpha == 0 else (1 - conv.alpha) * xv + conv.alpha * x0)


class HCCFDiffusionEncoder(HCCFEncoder):
    def __init__(self, conf, data):
        super().__init__(conf, data)
        del self.hgnnlayer
        args = LocalAwareEncoder.init_edhnn_config(None, self.latent_size)
        self.edhnn_user_n, self.edhnn_item_n = data.n_users + self.n_edges, data.n_items + self.n_edges
        self.edhnnlayer = _EquivSetGNNPattern(self.latent_size, args)

    def forward(self, keep_rate=0.5, device_rng=False):
        n_users = self.data.n_users
        emb = self.embedding_dict
        hidden = [torch.cat([emb['user_emb'], emb['item_emb']], 0)]
        gcn_hidden, hgnn_hidden = [], []
        with torch.no_grad():  # only the sign pattern of E W is used (nonzero(H > 0) passes no gradient, HCCF_diffusion.py:384)
            hyper_uu = ops.rows_times_small(emb['user_emb'].contiguous(), None, emb['user_w'].contiguous())
            hyper_ii = ops.rows_times_small(emb['item_emb'].contiguous(), None, emb['item_w'].contiguous())
        step = self._drop_step() if device_rng and keep_rate != 1.0 else None
        for layer in range(self.n_layers):
            cur = hidden[-1]
            gcn = self.gcnlayer(self.edgeDropper(self.sparse_norm_adj, keep_rate, None, self._drop_seed + layer if step is not None else None, step), cur)
            bu = (self.drop_out(hyper_uu) > 0).to(torch.float32)
            bi = (self.drop_out(hyper_ii) > 0).to(torch.float32)
            hyp = torch.cat([self.edhnnlayer(cur[:n_users], bu, self.edhnn_user_n), self.edhnnlayer(cur[n_users:], bi, self.edhnn_item_n)], 0)
            gcn_hidden.append(gcn)
            hgnn_hidden.append(hyp)
            hidden.append(gcn + hyp)
        out = sum(hidden)
        return out[:n_users], out[n_users:], gcn_hidden, hgnn_hidden


class EquivSetConvAttention(EquivSetConvScatter):
    """``forward(X, vertex, edges, atts, X0)`` of model/graph/HD2.py:624-643: ``atts [nnz, 1]`` scales ``W1(X)[vertex]`` before the
    mean over each hyperedge.  The weights become CSR values of the node -> hyperedge propagation (``ops.scatter_mean_conv_weighted``)."""

    def forward(self, X, vertex, edges, atts, X0):
        inc = self.incidence(vertex, edges, X.shape[-2])
        if getattr(self, "_pos_key", None) is not self._inc_key:
            self._pos, self._pos_key = graph.pair_positions(inc, vertex, edges), self._inc_key
        Xv = ops.scatter_mean_conv_weighted(inc, self.W1(X), atts, vertex, edges, self._pos)
        X = Xv if self.alpha == 0 else (1 - self.alpha) * Xv + self.alpha * X0
        return self.W(X)
