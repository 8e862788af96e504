"""Host-side mirror of the reference's graph encoders, running on libhgr.so.

Same class names, constructor arguments, ``forward`` signatures, return values and ``state_dict``
keys as the reference modules they replace, so a ``GraphRecommender`` subclass of the reference
keeps working when these are imported instead:

===========================  ==================================================================
``TorchGraphInterface``       base/torch_interface.py:3-19
``LGCN_Encoder``              model/graph/LightGCN.py:104-140
``SGL_Encoder``               model/graph/SGL.py:110-180 (+ data/augmentor.py through ``augmentor.py``)
``HGCNConv``                  model/graph/HGNN_HD3.py:540-553 (18 identical copies)
``SpAdjDropEdge``             model/graph/HCCF.py:213-226
``MLP``                       model/layers/MLP.py:29-117
``EquivSetConv``              model/graph/HGNN_HD3.py:655-720 (= model/layers/EquivSetConv.py:86-107)
``EquivSetGNN``               model/graph/HGNN_HD3.py:555-610
``LocalAwareEncoder``         model/graph/HGNN_HD3.py:352-427
``HCCFEncoder``               model/graph/HCCF.py:136-191 (+ GCNLayer :193-199, HGNNLayer :201-211)
``SHTEncoder``                model/graph/SHT.py:142-203
``DHCF_Encoder``              model/graph/DHCF.py:135-186
===========================  ==================================================================

What changes underneath: every ``torch.sparse.mm`` is ``ops.spmm`` / ``ops.hgconv`` /
``ops.lightgcn_propagate`` on a ``DeviceCSR``; LayerNorm + residual + activation + layer readout are
fused into the propagation kernel's epilogue.  The dense ``(U+I)^2`` copies of the adjacency that
the reference's ``LocalAwareEncoder`` keeps (``hyper_uu``/``hyper_ii``, HGNN_HD3.py:386-387, never
read by ``forward``) are not built.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import DeviceCSR


def _device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")


class TorchGraphInterface(object):
    """``convert_sparse_mat_to_tensor`` returns a ``DeviceCSR`` already resident in HBM, so the
    ``.cuda()`` / ``.to(device)`` the reference chains onto it are no-ops."""

    @staticmethod
    def convert_sparse_mat_to_tensor(X, device=None) -> DeviceCSR:
        if isinstance(X, DeviceCSR):
            return X
        return DeviceCSR.from_scipy(X, device=device or _device())

    @staticmethod
    def sparse_identity(n, device=None) -> DeviceCSR:
        dev = device or _device()
        return DeviceCSR(torch.arange(n + 1, dtype=torch.int64, device=dev), torch.arange(n, dtype=torch.int32, device=dev),
                         torch.ones(n, dtype=torch.float32, device=dev), (n, n), symmetric=True)


def _adjacency_of(data) -> DeviceCSR:
    """``data.norm_adj`` as a DeviceCSR (built once and cached on the data object)."""
    dev = getattr(data, "norm_adj_device", None)
    if dev is None:
        dev = TorchGraphInterface.convert_sparse_mat_to_tensor(data.norm_adj)
        try:
            data.norm_adj_device = dev
        except AttributeError:
            pass
    return dev


class LGCN_Encoder(nn.Module):
    def __init__(self, data, emb_size, n_layers):
        super(LGCN_Encoder, self).__init__()
        self.data = data
        self.latent_size = emb_size
        self.layers = n_layers
        self.norm_adj = data.norm_adj
        self.embedding_dict = self._init_model()
        self.sparse_norm_adj = _adjacency_of(data)

    def _init_model(self):
        initializer = nn.init.xavier_uniform_
        return nn.ParameterDict({
            'user_emb': nn.Parameter(initializer(torch.empty(self.data.n_users, self.latent_size))),
            'item_emb': nn.Parameter(initializer(torch.empty(self.data.n_items, self.latent_size))),
        })

    def forward(self):
        ego_embeddings = torch.cat([self.embedding_dict['user_emb'], self.embedding_dict['item_emb']], 0)
        all_embeddings = ops.lightgcn_propagate(self.sparse_norm_adj, ego_embeddings, self.layers)
        return all_embeddings[:self.data.n_users], all_embeddings[self.data.n_users:]


class SGL_Encoder(nn.Module):
    """``SGL_Encoder`` (model/graph/SGL.py:110-180): LightGCN propagation on the clean graph or on a perturbed one (a single
    ``DeviceCSR`` or one per layer), two augmented views per batch and ``InfoNCE`` between them.  ``graph_reconstruction`` reads the
    interaction list from ``data.dense_training_pairs()`` (the ``data.Interaction`` facade) or ``data.train_u / data.train_i``."""

    def __init__(self, data, emb_size, drop_rate, n_layers, temp, aug_type):
        super(SGL_Encoder, self).__init__()
        self.data = data
        self.drop_rate = drop_rate
        self.emb_size = emb_size
        self.n_layers = n_layers
        self.temp = temp
        self.aug_type = aug_type
        self.norm_adj = data.norm_adj
        initializer = nn.init.xavier_uniform_
        self.embedding_dict = nn.ParameterDict({
            'user_emb': nn.Parameter(initializer(torch.empty(self.data.n_users, self.emb_size))),
            'item_emb': nn.Parameter(initializer(torch.empty(self.data.n_items, self.emb_size))),
        })
        self.sparse_norm_adj = _adjacency_of(data)

    def graph_reconstruction(self):
        # the reference's condition `self.aug_type==0 or 1` is always true: one perturbed graph for all layers
        return self.random_graph_augment()

    def _training_pairs(self):
        """Distinct dense (user, item) pairs on the device: the reference drops nodes / edges of ``interaction_mat.nonzero()``
        (data/augmentor.py:12-42), i.e. of DISTINCT interactions with unit weights, whatever the file repeats."""
        pairs = getattr(self, "_pairs", None)
        if pairs is None:
            d = self.data
            u, i = d.dense_training_pairs() if hasattr(d, "dense_training_pairs") else (d.train_u, d.train_i)
            dev = self.embedding_dict['user_emb'].device
            key = torch.unique(torch.as_tensor(u).to(dev, torch.int64) * d.n_items + torch.as_tensor(i).to(dev, torch.int64))
            pairs = self._pairs = (torch.div(key, d.n_items, rounding_mode="floor").to(torch.int32), (key % d.n_items).to(torch.int32))
        return pairs

    def random_graph_augment(self):
        from . import augmentor

        u, i = self._training_pairs()
        return augmentor.random_graph_augment(u, i, self.data.n_users, self.data.n_items, self.aug_type, self.drop_rate)

    def forward(self, perturbed_adj=None):
        ego_embeddings = torch.cat([self.embedding_dict['user_emb'], self.embedding_dict['item_emb']], 0)
        if perturbed_adj is None or not isinstance(perturbed_adj, list):
            adj = self.sparse_norm_adj if perturbed_adj is None else perturbed_adj
            all_embeddings = ops.lightgcn_propagate(adj, ego_embeddings, self.n_layers)
        else:
            acc = ego_embeddings
            for k in range(self.n_layers):
                ego_embeddings = ops.spmm(perturbed_adj[k], ego_embeddings)
                acc = acc + ego_embeddings
            all_embeddings = acc / (self.n_layers + 1)
        return torch.split(all_embeddings, [self.data.n_users, self.data.n_items])

    def cal_cl_loss(self, idx, perturbed_mat1, perturbed_mat2):
        from .loss_torch import InfoNCE

        dev = self.embedding_dict['user_emb'].device
        u_idx = torch.unique(torch.as_tensor(idx[0], device=dev).long())
        i_idx = torch.unique(torch.as_tensor(idx[1], device=dev).long())
        user_view_1, item_view_1 = self.forward(perturbed_mat1)
        user_view_2, item_view_2 = self.forward(perturbed_mat2)
        view1 = torch.cat((user_view_1[u_idx], item_view_1[i_idx]), 0)
        view2 = torch.cat((user_view_2[u_idx], item_view_2[i_idx]), 0)
        return InfoNCE(view1, view2, self.temp)


class HGCNConv(nn.Module):
    def __init__(self, leaky):
        super(HGCNConv, self).__init__()
        self.leaky = float(leaky)
        self.act = nn.LeakyReLU(negative_slope=leaky)

    def forward(self, adj, embs, act=True, ln=None, residual=None):
        """``act(adj @ (adj.t() @ embs))``; ``ln`` (an ``nn.LayerNorm``) and ``residual`` fuse the
        ``lns[k](...) + res`` the reference applies right after every call."""
        return ops.hgconv(adj, embs, self.leaky if act else None,
                          None if ln is None else ln.weight, None if ln is None else ln.bias, residual,
                          1e-5 if ln is None else ln.eps)


def drop_edges(adj: DeviceCSR, keep: float, mask: torch.Tensor | None = None, seed: int | None = None,
               step: torch.Tensor | None = None) -> DeviceCSR:
    """``hgr.drop_edges(adj, keep, mask)`` (SURVEY.md 8b seam 6) on ``hgr_drop_edges_f32``: the matrix with THIS sparsity
    pattern whose entries are kept with probability ``keep`` and rescaled by ``1 / keep`` (dropped ones become explicit zeros), plus
    the same for its transpose (what the backward propagation reads) -- two launches, no sort, no compaction.
    ``mask``: uniform numbers of ``torch.rand(nnz)`` (any device) to replay the reference's random stream, entry for entry;
    ``None``: Philox on the device keyed by ``seed`` and the entry's coordinates; ``step`` (a one-element int64 CUDA tensor) is mixed
    into the seed when the kernel RUNS, so a captured CUDA graph draws a new mask at every replay."""
    import ctypes as C

    from . import _lib

    lib = _lib.lib()
    nnz = adj._nnz()
    dev = adj.device
    vals = torch.empty(nnz, dtype=torch.float32, device=dev)
    rand = None if mask is None else mask.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    sd = 0 if seed is None else int(seed) & ((1 << 64) - 1)

    def run(mirror, pos, out):
        _lib.check(lib.hgr_drop_edges_f32(adj.indptr.data_ptr(), adj.indices.data_ptr(), adj.values.data_ptr(), adj.shape[0], adj.shape[1],
                                          nnz, float(keep), C.c_uint64(sd), _lib.ptr(step), _lib.ptr(rand), _lib.ptr(pos), mirror,
                                          out.data_ptr(), _lib.stream_ptr()))

    run(0, None, vals)
    if not (adj.symmetric or getattr(adj, "_tperm", None) is not None) or adj.shape[0] != adj.shape[1]:
        return adj.with_values(vals)  # the transpose of a general matrix is built on demand (DeviceCSR.t)
    t_vals = torch.empty(nnz, dtype=torch.float32, device=dev)
    run(1, adj.transpose_permutation() if rand is not None else None, t_vals)
    return adj.with_values(vals, t_vals)


class SpAdjDropEdge(nn.Module):
    """Bernoulli edge dropout (``drop_edges``).  Default: the reference's CPU random stream (``torch.rand(nnz)`` on the host
    generator, HCCF.py:224), so a seeded run keeps and rescales exactly the same edges; ``rand`` = uniform numbers drawn elsewhere;
    ``seed`` = draw on the device with Philox instead (no host work, no upload)."""

    def __init__(self):
        super(SpAdjDropEdge, self).__init__()

    def forward(self, adj: DeviceCSR, keepRate, rand=None, seed=None, step=None):
        if keepRate == 1.0:
            return adj
        if rand is None and seed is None:
            rand = torch.rand(adj._nnz())
        if adj.shape[0] == adj.shape[1] and (adj.symmetric or getattr(adj, "_tperm", None) is not None):
            return drop_edges(adj, keepRate, rand, seed, step)
        # rectangular / asymmetric matrices: compacted like the reference does (their transpose is rebuilt by DeviceCSR.t)
        if rand is None:
            rand = torch.rand(adj._nnz(), device=adj.device)
        mask = ((rand.to(adj.device) + keepRate).floor()).type(torch.bool)
        keep = torch.full((), float(keepRate), dtype=torch.float32, device=adj.device)
        rows = torch.repeat_interleave(torch.arange(adj.shape[0], device=adj.device), adj.indptr[1:] - adj.indptr[:-1])
        counts = torch.bincount(rows[mask], minlength=adj.shape[0])
        indptr = torch.zeros(adj.shape[0] + 1, dtype=torch.int64, device=adj.device)
        torch.cumsum(counts, 0, out=indptr[1:])
        return DeviceCSR(indptr, adj.indices[mask], adj.values[mask] / keep, adj.shape, chunk_nnz=adj.chunk_nnz, split=adj.split)


class MLP(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout=.5, Normalization='bn', InputNorm=False):
        super(MLP, self).__init__()
        assert Normalization in ['bn', 'ln', 'None']
        norm = {'bn': nn.BatchNorm1d, 'ln': nn.LayerNorm, 'None': lambda c: nn.Identity()}[Normalization]
        self.in_channels, self.hidden_channels, self.out_channels = in_channels, hidden_channels, out_channels
        self.lins = nn.ModuleList()
        self.normalizations = nn.ModuleList()
        self.InputNorm = InputNorm
        self.normalizations.append(norm(in_channels) if InputNorm and Normalization != 'None' else nn.Identity())
        if num_layers == 1:
            self.lins.append(nn.Linear(in_channels, out_channels))
        else:
            self.lins.append(nn.Linear(in_channels, hidden_channels))
            self.normalizations.append(norm(hidden_channels))
            for _ in range(num_layers - 2):
                self.lins.append(nn.Linear(hidden_channels, hidden_channels))
                self.normalizations.append(norm(hidden_channels))
            self.lins.append(nn.Linear(hidden_channels, out_channels))
        self.dropout = dropout

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()
        for normalization in self.normalizations:
            if not (normalization.__class__.__name__ == 'Identity'):
                normalization.reset_parameters()

    @staticmethod
    def _norm(module, x):
        # LayerNorm over rows of 32 / 64 / 128 floats runs on libhgr's row kernel (torch's takes 2.2 ms for 1.5 M x 64)
        if (isinstance(module, nn.LayerNorm) and module.elementwise_affine and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
                and x.shape[1] in (32, 64, 128) and len(module.normalized_shape) == 1):
            return ops.layer_norm(x, module.weight, module.bias, module.eps)
        return module(x)

    def forward(self, x):
        x = self._norm(self.normalizations[0], x)
        for i, lin in enumerate(self.lins[:-1]):
            x = ops.linear(x, lin.weight, lin.bias, relu=True)
            x = self._norm(self.normalizations[i + 1], x)
            x = F.dropout(x, p=self.dropout, training=self.training)
        return ops.linear(x, self.lins[-1].weight, self.lins[-1].bias)


class EquivSetConv(nn.Module):
    def __init__(self, in_features, out_features, ncount, mcount, mlp1_layers=1, mlp2_layers=1, mlp3_layers=1, aggr='add',
                 alpha=0.5, dropout=0., normalization='None', input_norm=False, hypergraph=None, data=None):
        super().__init__()
        self.in_features = in_features
        self.W1 = MLP(in_features, out_features, out_features, mlp1_layers, dropout=dropout, Normalization=normalization,
                      InputNorm=input_norm) if mlp1_layers > 0 else nn.Identity()
        self.W2 = MLP(in_features + out_features, out_features, out_features, mlp2_layers, dropout=dropout,
                      Normalization=normalization, InputNorm=input_norm) if mlp2_layers > 0 else None
        self.W = MLP(out_features, out_features, out_features, mlp3_layers, dropout=dropout, Normalization=normalization,
                     InputNorm=input_norm) if mlp3_layers > 0 else nn.Identity()
        self.aggr = aggr
        self.alpha = alpha
        self.dropout = dropout
        self.data = data
        self.hgcn_layers = nn.ModuleList([HGCNConv(0.5) for i in range(2)])
        self.mean_pooling = nn.AdaptiveAvgPool1d(out_features)
        self.lns = torch.nn.ModuleList([torch.nn.LayerNorm(out_features) for i in range(2)])

    def reset_parameters(self):
        for m in (self.W1, self.W2, self.W):
            if isinstance(m, MLP):
                m.reset_parameters()

    def forward(self, X, sparse_norm_adj, X0, ui_adj=None, act=True):
        Xve = self.W1(X)
        # Xe = lns[0](hgcn(A, Xve)) + Xve : one fused two-stage propagation
        Xe = self.hgcn_layers[0](sparse_norm_adj, Xve, act=True, ln=self.lns[0], residual=Xve)
        if self.W2 is None:
            Xev = Xe  # the reference slices cat([X, Xe])[..., in_features:], which is Xe
        else:
            Xev = self.W2(torch.cat([X, Xe], -1))
        if Xev.shape[-1] != self.mean_pooling.output_size:
            Xev = self.mean_pooling(Xev)  # identity when the width already matches
        X_v = self.hgcn_layers[1](sparse_norm_adj, Xev, act=True, ln=self.lns[1], residual=Xev)
        X = X_v
        if self.alpha != 0:
            X = (1 - self.alpha) * X + self.alpha * X0
        return self.W(X)


class EquivSetConvScatter(nn.Module):
    """``EquivSetConv`` of the scatter form (model/layers/layers2/EquivSetConv2.py:37-100, = layers3/EquivSetConv3.py,
    HCCF_diffusion.py:257-308): ``W1`` -> mean over each hyperedge -> ``W2`` -> mean over each vertex ->
    restart mix -> ``W``.  The two ``torch_scatter.scatter(..., reduce='mean')`` calls become deterministic
    propagations over a row-normalised incidence pair built once per (vertex, edges) index pair.  ``W2`` with
    ``mlp2_layers == 0`` is the reference's slice (``Xev = Xe[E]``), the only form its configs use."""

    def __init__(self, in_features, out_features, mlp1_layers=1, mlp2_layers=1, mlp3_layers=1, aggr='add', alpha=0.5,
                 dropout=0., normalization='None', input_norm=False, hypergraph=None, data=None):
        super().__init__()
        if aggr != 'mean':
            raise NotImplementedError("only reduce='mean' (the reference's configs) is on the B200 path")
        if mlp2_layers > 0:
            raise NotImplementedError("W2 as an MLP over [X[V], Xe[E]] pairs is not used by any shipped config")
        self.W1 = MLP(in_features, out_features, out_features, mlp1_layers, dropout=dropout, Normalization=normalization,
                      InputNorm=input_norm) if mlp1_layers > 0 else nn.Identity()
        self.W = MLP(out_features, out_features, out_features, mlp3_layers, dropout=dropout, Normalization=normalization,
                     InputNorm=input_norm) if mlp3_layers > 0 else nn.Identity()
        self.aggr, self.alpha, self.dropout, self.data = aggr, alpha, dropout, data
        self._inc_key, self._inc = None, None

    def reset_parameters(self):
        for m in (self.W1, self.W):
            if isinstance(m, MLP):
                m.reset_parameters()

    def incidence(self, vertex, edges, n_nodes):
        from .graph import build_incidence

        key = (vertex.data_ptr(), edges.data_ptr(), int(vertex.numel()), int(n_nodes))
        if key != self._inc_key:
            self._inc, self._inc_key = build_incidence(vertex, edges, n_nodes, device=_device()), key
        return self._inc

    def forward(self, X, vertex, edges, X0):
        return self.forward_inc(X, self.incidence(vertex, edges, X.shape[-2]), X0)

    def forward_inc(self, X, inc, X0):
        """The same convolution on an incidence pair the caller built once (``graph.Incidence``)."""
        Xv = ops.segment_mean_to_nodes(inc, ops.segment_mean_to_edges(inc, self.W1(X)))
        X = Xv if self.alpha == 0 else (1 - self.alpha) * Xv + self.alpha * X0
        return self.W(X)


class EquivSetGNN(nn.Module):
    def __init__(self, num_features, args, dense_hypergraph, data, ncount, mcount):
        super().__init__()
        act = {'Id': nn.Identity(), 'relu': nn.ReLU(), 'prelu': nn.PReLU()}
        self.act = act[args['activation']]
        self.input_drop = nn.Dropout(args['input_dropout'])
        self.dropout = nn.Dropout(args['dropout'])
        self.data = data
        self.in_channels = num_features
        self.hidden_channels = args['MLP_hidden']
        self.mlp1_layers = args['MLP_num_layers']
        self.mlp2_layers = args['MLP_num_layers'] if args['MLP2_num_layers'] < 0 else args['MLP2_num_layers']
        self.mlp3_layers = args['MLP_num_layers'] if args['MLP3_num_layers'] < 0 else args['MLP3_num_layers']
        self.nlayer = args['All_num_layers']
        self.lin_in = torch.nn.Linear(num_features, args['MLP_hidden'])
        self.conv = EquivSetConv(args['MLP_hidden'], args['MLP_hidden'], ncount, mcount, mlp1_layers=self.mlp1_layers,
                                 mlp2_layers=self.mlp2_layers, mlp3_layers=self.mlp3_layers, alpha=args['restart_alpha'],
                                 aggr=args['aggregate'], dropout=args['dropout'], normalization=args['normalization'],
                                 input_norm=args['AllSet_input_norm'], hypergraph=dense_hypergraph, data=self.data)

    def reset_parameters(self):
        self.lin_in.reset_parameters()
        self.conv.reset_parameters()

    def forward(self, x, sparse_norm_adj, n_nodes=None, ui_adj=None, act=True):
        x = self.dropout(x)
        # (sharded graph: the rows of this Linear are the input of the first propagation -- its epilogue publishes them)
        x = ops.linear(x, self.lin_in.weight, self.lin_in.bias, relu=True,
                       publish=sparse_norm_adj if ops._sharded(sparse_norm_adj) and not (self.training and self.dropout.p > 0) else None)
        x0 = x
        for i in range(self.nlayer):
            x = self.dropout(x)
            x = self.conv(x, sparse_norm_adj, x0, ui_adj, act)
            x = self.act(x)
        x = self.dropout(x)
        return x


class LocalAwareEncoder(nn.Module):
    def __init__(self, data, emb_size, hyper_size, n_layers, leaky, drop_rate, device=None, use_self_att=False):
        super(LocalAwareEncoder, self).__init__()
        self.data = data
        self.latent_size = emb_size
        self.hyper_size = hyper_size
        self.layers = n_layers
        self.norm_adj = data.norm_adj
        self.relu = nn.ReLU()
        self.act = nn.LeakyReLU(leaky)
        self.dropout = nn.Dropout(drop_rate)
        self.edgeDropper = SpAdjDropEdge()
        self.sparse_norm_adj = _adjacency_of(data)
        self.edhnn_args = self.init_edhnn_config(self.hyper_size)
        self.hgcn_layer = HGCNConv(leaky=0.3)
        self.hgnn_layers = nn.ModuleList([HGCNConv(leaky=0.3) for i in range(self.layers)])
        self.edhnn_layers = nn.ModuleList([EquivSetGNN(hyper_size, self.edhnn_args, self.sparse_norm_adj, self.data,
                                                       self.data.n_users, self.data.n_items) for i in range(self.layers)])
        self.lns = torch.nn.ModuleList([torch.nn.LayerNorm(hyper_size) for i in range(self.layers)])
        self.edhnn_ui_n = self.data.n_users + self.data.n_items

    def init_edhnn_config(self, hyper_size):
        return {'MLP_hidden': hyper_size, 'MLP1_num_layers': 0, 'MLP2_num_layers': 0, 'MLP3_num_layers': 1, 'MLP_num_layers': 0,
                'restart_alpha': 0.0, 'aggregate': 'mean', 'dropout': 0.5, 'normalization': 'ln', 'input_norm': True,
                'All_num_layers': 1, 'activation': 'relu', 'input_dropout': 0.6, 'AllSet_input_norm': True}

    def forward(self, ego_embeddings, sparse_norm_adj):
        res = ego_embeddings
        for k in range(self.layers):
            if k != self.layers - 1:
                # "+ res" on hgr_add_rows_f32; when the graph is sharded the sum reaches every rank's table from the same kernel
                ego_embeddings = ops.add_rows(self.edhnn_layers[k](ego_embeddings, sparse_norm_adj, self.edhnn_ui_n, None), res,
                                              publish=self.sparse_norm_adj if ops._sharded(self.sparse_norm_adj) else None)
            else:
                # lns[k](hgcn(A, ego, act=False)) + res in one fused two-stage propagation; like the
                # reference this last layer always uses the un-dropped adjacency
                ego_embeddings = self.hgcn_layer(self.sparse_norm_adj, ego_embeddings, act=False, ln=self.lns[k], residual=res)
        return ego_embeddings[:self.data.n_users], ego_embeddings[self.data.n_users:]


class GCNLayer(nn.Module):
    def __init__(self, leaky):
        super(GCNLayer, self).__init__()
        self.act = nn.LeakyReLU(negative_slope=leaky)

    def forward(self, adj, embeds):
        return ops.spmm(adj, embeds)


class HGNNLayer(nn.Module):
    def __init__(self, leaky):
        super(HGNNLayer, self).__init__()
        self.act = nn.LeakyReLU(negative_slope=leaky)

    def forward(self, adj, embeds):
        # adj [n, hyper_dim] dense learned incidence, embeds [n, D]: two tall-and-skinny products on libhgr (csrc/hyperedge.cu)
        return ops.hyperedge(adj, embeds)  # raises for widths the kernels do not cover: there is no library fallback


def _dense_incidence(emb, w):
    """``emb @ w`` (HCCF.py:178-179): tall-and-skinny in the forward pass, and its weight gradient ``emb.T @ dH`` contracts
    over all nodes -- libhgr kernels (widths 32 / 64 / 128; anything else raises, there is no library fallback)."""
    return ops.tall_times_small(emb, w)


class HCCFEncoder(nn.Module):
    def __init__(self, conf, data):
        super(HCCFEncoder, self).__init__()
        self.data = data
        self._parse_config(conf)
        self.gcnlayer = GCNLayer(self.leaky)
        self.hgnnlayer = HGNNLayer(self.leaky)
        self.norm_adj = data.norm_adj
        self.sparse_norm_adj = _adjacency_of(data)
        self.embedding_dict = self._init_model()
        self.drop_out = nn.Dropout(self.drop_rate)
        self.edgeDropper = SpAdjDropEdge()

    def _parse_config(self, config):
        self.lRate = float(config['lrate'])
        self.lr_decay = float(config['lr_decay'])
        self.maxEpoch = int(config['max_epoch'])
        self.batchSize = int(config['batch_size'])
        self.reg = float(config['reg'])
        self.latent_size = int(config['embedding_size'])
        self.hyperDim = int(config['hyper_dim'])
        self.drop_rate = float(config['drop_rate'])
        self.leaky = float(config['p'])
        self.n_layers = int(config['n_layers'])
        self.n_edges = int(config['hyper_dim'])

    def _init_model(self):
        initializer = nn.init.xavier_uniform_
        return nn.ParameterDict({
            'user_emb': nn.Parameter(initializer(torch.empty(self.data.n_users, self.latent_size))),
            'item_emb': nn.Parameter(initializer(torch.empty(self.data.n_items, self.latent_size))),
            'user_w': nn.Parameter(initializer(torch.empty(self.latent_size, self.n_edges))),
            'item_w': nn.Parameter(initializer(torch.empty(self.latent_size, self.n_edges))),
        })

    def _drop_step(self):
        """Device-side step counter mixed into the Philox seed of the edge-drop kernel (advanced once per forward by a captured
        add, so a replayed CUDA graph draws new masks every step)."""
        ctr = getattr(self, "_drop_ctr", None)
        if ctr is None:
            ctr = self._drop_ctr = torch.zeros(1, dtype=torch.int64, device=self.sparse_norm_adj.device)
            self._drop_seed = int(torch.initial_seed()) & 0x7fffffff
        ctr += 1
        return ctr

    def forward(self, keep_rate=0.5, device_rng=False):
        """``device_rng``: draw each layer's edge-drop mask with Philox inside ``hgr_drop_edges_f32`` instead of the reference's
        CPU ``torch.rand(nnz)`` + upload (HCCF.py:217-226); the default replays the reference's CPU random stream."""
        n_users = self.data.n_users
        embeddings = torch.cat([self.embedding_dict['user_emb'], self.embedding_dict['item_emb']], 0)
        hidden = [embeddings]
        gcn_hidden = []
        hgnn_hidden = []
        hyper_uu = _dense_incidence(self.embedding_dict['user_emb'], self.embedding_dict['user_w'])
        hyper_ii = _dense_incidence(self.embedding_dict['item_emb'], self.embedding_dict['item_w'])
        step = self._drop_step() if device_rng and keep_rate != 1.0 else None
        for i in range(self.n_layers):
            gcn_emb = self.gcnlayer(self.edgeDropper(self.sparse_norm_adj, keep_rate, None, self._drop_seed + i if step is not None else None, step),
                                    hidden[-1])
            hyper_uemb = self.hgnnlayer(self.drop_out(hyper_uu), hidden[-1][:n_users])
            hyper_iemb = self.hgnnlayer(self.drop_out(hyper_ii), hidden[-1][n_users:])
            gcn_hidden += [gcn_emb]
            hgnn_hidden += [torch.cat([hyper_uemb, hyper_iemb], 0)]
            hidden += [gcn_emb + hgnn_hidden[-1]]
        embeddings = sum(hidden)
        return embeddings[:n_users], embeddings[n_users:], gcn_hidden, hgnn_hidden


class SHTEncoder(nn.Module):
    """``SHTEncoder`` (model/graph/SHT.py:142-203): LightGCN propagation with SUM readout over the normalised adjacency, then the
    low-rank hypergraph transform ``embeds @ (hyper.T @ hyper)`` on the detached user and item halves.  Same constructor, parameter
    names (``uEmbeds``, ``iEmbeds``, ``uHyper``, ``iHyper``) and return triple as the reference."""

    def __init__(self, data, args):
        super(SHTEncoder, self).__init__()
        init = nn.init.xavier_uniform_
        self.args = args
        self.data = data
        self.n_user = self.data.n_users
        self.n_item = self.data.n_items
        self._parse_config(args)
        self.norm_adj = self.data.norm_adj
        self.sparse_norm_adj = _adjacency_of(data)
        self.uEmbeds = nn.Parameter(init(torch.empty(self.n_user, self.embeddingSize)))
        self.iEmbeds = nn.Parameter(init(torch.empty(self.n_item, self.embeddingSize)))
        self.uHyper = nn.Parameter(init(torch.empty(args['hyperedge_num'], self.embeddingSize)))
        self.iHyper = nn.Parameter(init(torch.empty(args['hyperedge_num'], self.embeddingSize)))

    def _parse_config(self, kwargs):
        self.maxEpoch = int(kwargs['max_epoch'])
        self.batchSize = int(kwargs['batch_size'])
        self.lRate = float(kwargs['lrate'])
        self.lr_decay = float(kwargs['lr_decay'])
        self.reg = float(kwargs['reg'])
        self.latent_size = int(kwargs['embedding_size'])
        self.embeddingSize = int(kwargs['hyper_dim'])
        self.hyperDim = int(kwargs['hyper_dim'])
        self.dropRate = float(kwargs['drop_rate'])
        self.negSlove = float(kwargs['p'])
        self.nLayers = int(kwargs['n_layers'])
        self.ss_rate = float(kwargs['cl_rate'])
        self.temp = float(kwargs['temp'])
        self.seed = int(kwargs['seed'])
        self.edgeSampRate = 0.1
        self.ssl1_reg = 0.1
        self.ssl2_reg = 0.1
        self.early_stopping_steps = int(kwargs['early_stopping_steps'])
        self.hyperedge_num = int(kwargs['hyperedge_num'])

    def gcnLayer(self, adj, embeds):
        return ops.spmm(adj, embeds)

    def hgnnLayer(self, embeds, hyper):
        small = hyper.T @ hyper  # [D, D]: a D x D product of two parameter matrices, not a pass over the nodes
        return ops.tall_times_small(embeds, small)

    def forward(self):
        embeds = torch.concat([self.uEmbeds, self.iEmbeds], dim=0)
        # lats = [E, A E, A^2 E, ...]; embeds = sum(lats): the fused propagation with the sum readout (SHT.py:192-199)
        embeds = ops.lightgcn_propagate(self.sparse_norm_adj, embeds, self.nLayers, sum_readout=True)
        # this detach helps eliminate the mutual influence between the local GCN and the global HGNN (SHT.py:200)
        hyperUEmbeds = self.hgnnLayer(embeds[:self.n_user].detach(), self.uHyper)
        hyperIEmbeds = self.hgnnLayer(embeds[self.n_user:].detach(), self.iHyper)
        return embeds, hyperUEmbeds, hyperIEmbeds


class DHCF_Encoder(nn.Module):
    """``DHCF_Encoder`` (model/graph/DHCF.py:135-186): ``HGCNConv`` on the RECTANGULAR user x item interaction matrix --
    ``leaky(R (R^T U))`` for users, ``leaky(R^T (R I))`` for items, every layer applied to the INPUT tables as in the reference --
    and a concatenation readout.  The reference densifies ``R`` (``.to_dense()``, U x I floats); here it stays a ``DeviceCSR``."""

    def __init__(self, config, data, args):
        super(DHCF_Encoder, self).__init__()
        self.data = data
        self.adj = TorchGraphInterface.convert_sparse_mat_to_tensor(data.interaction_mat)
        self._parse_args(args)
        self.embedding_dict = self._init_model()
        self.fc_u = nn.Linear(self.hyper_dim, self.hyper_dim)
        self.fc_i = nn.Linear(self.hyper_dim, self.hyper_dim)
        self.hgnn_u = [HGCNConv(leaky=self.p) for _ in range(self.layers)]
        self.hgnn_i = [HGCNConv(leaky=self.p) for _ in range(self.layers)]
        self.non_linear = nn.ReLU()
        self.dropout = nn.Dropout(self.drop_rate)

    def _parse_args(self, args):
        self.input_dim = args['input_dim']
        self.hyper_dim = args['hyper_dim']
        self.p = args['p']
        self.drop_rate = args['drop_rate']
        self.layers = args['n_layers']

    def _init_model(self):
        initializer = nn.init.xavier_uniform_
        return nn.ParameterDict({
            'user_emb': nn.Parameter(initializer(torch.empty(self.data.n_users, self.hyper_dim))),
            'item_emb': nn.Parameter(initializer(torch.empty(self.data.n_items, self.hyper_dim))),
        })

    def forward(self):
        uEmbed = self.embedding_dict['user_emb']
        iEmbed = self.embedding_dict['item_emb']
        user_embeds = [uEmbed]
        item_embeds = [iEmbed]
        for idx in range(self.layers):
            user_embeds.append(self.hgnn_u[idx](self.adj, uEmbed))
            item_embeds.append(self.hgnn_i[idx](self.adj.t(), iEmbed))
        return torch.cat(user_embeds, dim=1), torch.cat(item_embeds, dim=1)


class HGNNModel(nn.Module):
    """``HGNNModel`` of the hypergraph-diffusion recommender in ``--mode=local_only``
    (model/graph/HGNN_HD3.py:248-330): the two embedding tables plus ``LocalAwareEncoder``.  The
    group branch (``GroupAwareEncoder``: HWNN wavelet layers on dense N x N matrices) is out of scope
    (SURVEY.md 2.3)."""

    def __init__(self, data, args, device=None):
        super(HGNNModel, self).__init__()
        self.data = data
        self.device = device
        self.hyper_dim = int(args['hyper_dim'])
        self.layers = int(args['n_layers'])
        self.p = float(args.get('p', 0.3))
        self.drop_rate = float(args.get('drop_rate', 0.2))
        self.batchSize = int(args.get('batch_size', 2048))
        self.sparse_norm_adj = _adjacency_of(data)
        initializer = nn.init.xavier_uniform_
        self.embedding_dict = nn.ParameterDict({
            'user_emb': nn.Parameter(initializer(torch.empty(self.data.n_users, self.hyper_dim))),
            'item_emb': nn.Parameter(initializer(torch.empty(self.data.n_items, self.hyper_dim))),
        })
        self.hgnn_layer_local = LocalAwareEncoder(self.data, self.hyper_dim, self.hyper_dim, self.layers, self.p, self.drop_rate, device)
        self.edgeDropper = SpAdjDropEdge()

    def calculate_local_embeddings(self, keep_rate: float = 1):
        ego_embeddings = torch.cat([self.embedding_dict['user_emb'], self.embedding_dict['item_emb']], 0)
        sparse_norm_adj = self.edgeDropper(self.sparse_norm_adj, keep_rate)
        return self.hgnn_layer_local(ego_embeddings, sparse_norm_adj)

    def forward(self, mode='local', keep_rate=1):
        if mode != 'local':
            raise NotImplementedError("only the local (hypergraph-diffusion) branch is on the B200 path")
        return self.calculate_local_embeddings(keep_rate=keep_rate)
