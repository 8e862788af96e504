"""Propagation operators over ``DeviceCSR`` with hand-written backward passes.

Each function replaces one library call site of the reference (paths relative to HD_SELFRec/):

=====================  =====================================================================
``spmm(adj, X)``        ``torch.sparse.mm(adj, X)`` -- model/graph/LightGCN.py:133, HCCF.py:199
``hgconv(...)``         ``HGCNConv.forward`` -- model/graph/HGNN_HD3.py:540-553 (+17 copies), with the
                        ``lns[k](...) + res`` that always follows it fused in (HGNN_HD3.py:421,710,714)
``lightgcn_propagate``  ``LGCN_Encoder.forward`` body -- model/graph/LightGCN.py:131-136
=====================  =====================================================================

All of them run on the caller's current CUDA stream and raise if libhgr.so is missing.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .graph import DeviceCSR

_SUPPORTED_D = (32, 64, 128)

# bench.py sets this to a list to time the propagation calls live: every C-ABI propagation call then
# appends (start_event, end_event, number of SpMM launches inside the call).
PROFILE_EVENTS = None


class _Timed:
    def __init__(self, n_spmm):
        self.n = n_spmm

    def __enter__(self):
        if PROFILE_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if PROFILE_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE_EVENTS.append((self.e0, e1, self.n))


def _check_dense(x: torch.Tensor, rows: int, what: str) -> torch.Tensor:
    if not x.is_cuda:
        raise _lib.HgrError("%s must be a CUDA tensor (no CPU path)" % what)
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError("%s must be a 2-D float32 tensor, got %s %s" % (what, x.dtype, tuple(x.shape)))
    if x.shape[0] != rows:
        raise ValueError("%s has %d rows, the graph needs %d" % (what, x.shape[0], rows))
    if x.shape[1] not in _SUPPORTED_D:
        raise ValueError("embedding width %d unsupported (need one of %s)" % (x.shape[1], _SUPPORTED_D))
    return x.contiguous()


def _epilogue(slope=None, gamma=None, beta=None, eps=1e-5, residual=None, addends=(), scale=1.0, scale_always=False,
              pre=None, gather_ptrs=(), gather_row_offset=0, gather_mc=0) -> _lib.Epilogue:
    ep = _lib.Epilogue()
    ep.use_leaky = 0 if slope is None else 1
    ep.leaky_slope = 0.0 if slope is None else float(slope)
    ep.ln_gamma, ep.ln_beta, ep.ln_eps = _lib.ptr(gamma), _lib.ptr(beta), float(eps)
    ep.residual = _lib.ptr(residual)
    ep.n_addends = len(addends)
    for j, a in enumerate(addends):
        ep.addends[j] = a.data_ptr()
    ep.scale, ep.scale_always = float(scale), int(bool(scale_always))
    ep.pre = _lib.ptr(pre)
    ep.n_gather = len(gather_ptrs)  # fused all-gather: raw (peer-mapped) device pointers of every rank's gathered table
    for j, p in enumerate(gather_ptrs):
        ep.gather_out[j] = int(p)
    ep.gather_row_offset = int(gather_row_offset)
    ep.gather_mc = int(gather_mc) or None  # NVSwitch multicast address of the gathered table (one multimem.st instead of world stores)
    return ep


make_epilogue = _epilogue


def publish_rows(x: torch.Tensor, gather_ptrs, gather_row_offset: int, gather_mc: int = 0) -> None:
    """``hgr_publish_rows_f32``: store the owned rows ``x`` at ``gather_row_offset`` of every (peer-mapped) gathered table."""
    x = x.contiguous()
    g = _lib.make_gather(gather_ptrs, gather_row_offset, gather_mc)
    _lib.check(_lib.lib().hgr_publish_rows_f32(x.data_ptr(), x.shape[0], x.shape[1], C.byref(g), _lib.stream_ptr()))


def leaky_ln_bwd(pre, dy, gamma, eps, slope, gather_ptrs=(), gather_row_offset=0, gather_mc=0):
    """Backward of ``LayerNorm(leaky_relu(pre)) * gamma + beta`` (hgr_leaky_ln_bwd_f32): returns
    ``(dpre, dgamma, dbeta)``; ``gamma is None`` means no LayerNorm.  ``gather_ptrs``: also publish the ``dpre`` rows into
    every rank's gathered table (hgr_leaky_ln_bwd_gather_f32)."""
    d = dy.shape[1]
    dz = torch.empty_like(dy)
    dgamma = dbeta = parts = None
    if gamma is not None:
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        parts = torch.empty((_lib.lib().hgr_ln_bwd_partial_rows(dy.shape[0]), 2, d), dtype=torch.float32, device=dy.device)
    g = _lib.make_gather(gather_ptrs, gather_row_offset, gather_mc)
    _lib.check(_lib.lib().hgr_leaky_ln_bwd_gather_f32(pre.data_ptr(), dy.data_ptr(), _lib.ptr(gamma), float(eps), 0 if slope is None else 1,
                                                      0.0 if slope is None else float(slope), dy.shape[0], d, dz.data_ptr(),
                                                      _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(parts), C.byref(g), _lib.stream_ptr()))
    return dz, dgamma, dbeta


def _sharded(adj) -> bool:
    """A ``dist.DistGraph`` (rows partitioned across ranks) instead of a whole-matrix ``DeviceCSR``."""
    return getattr(adj, "world", 0) >= 1 and hasattr(adj, "part")


def _ws(a: DeviceCSR, d: int):
    ws = a.workspace(d)
    return (None, 0) if ws is None else (ws.data_ptr(), ws.numel() * 4)


def spmm_raw(a: DeviceCSR, x: torch.Tensor, ep: _lib.Epilogue | None = None, out: torch.Tensor | None = None,
             no_local_out: bool = False):
    """One ``hgr_spmm_f32`` call; no autograd.  ``no_local_out``: the epilogue's fused all-gather is the only
    destination (returns None)."""
    x = _check_dense(x, a.shape[1], "X")
    d = x.shape[1]
    if no_local_out:
        y = None
    else:
        y = out if out is not None else torch.empty((a.shape[0], d), dtype=torch.float32, device=x.device)
    ws, ws_bytes = _ws(a, d)
    with _Timed(1):
        _lib.check(_lib.lib().hgr_spmm_f32(C.byref(a.desc), x.data_ptr(), _lib.ptr(y), d, None if ep is None else C.byref(ep),
                                           ws, ws_bytes, _lib.stream_ptr()))
    return y


def hgconv_raw(a: DeviceCSR, at: DeviceCSR, x: torch.Tensor, ep: _lib.Epilogue | None = None) -> torch.Tensor:
    """One ``hgr_hgconv_f32`` call: epilogue(A (At X)); no autograd."""
    x = _check_dense(x, at.shape[1], "X")
    d = x.shape[1]
    tmp = torch.empty((at.shape[0], d), dtype=torch.float32, device=x.device)
    y = torch.empty((a.shape[0], d), dtype=torch.float32, device=x.device)
    w1, b1 = _ws(a, d)
    w2, b2 = _ws(at, d)
    ws, ws_bytes = (w1, b1) if b1 >= b2 else (w2, b2)
    with _Timed(2):
        _lib.check(_lib.lib().hgr_hgconv_f32(C.byref(a.desc), C.byref(at.desc), x.data_ptr(), tmp.data_ptr(), y.data_ptr(), d,
                                             None if ep is None else C.byref(ep), ws, ws_bytes, _lib.stream_ptr()))
    return y


class _Spmm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a):
        ctx.a = a
        return spmm_raw(a, x)

    @staticmethod
    def backward(ctx, dy):
        return spmm_raw(ctx.a.t(), dy.contiguous()), None


def spmm(adj: DeviceCSR, x: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``torch.sparse.mm(adj, x)``; backward is ``adj.t() @ dy`` through the same kernel."""
    if _sharded(adj):
        return adj.spmm(x)
    return _Spmm.apply(x, adj)


class _HGConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, a, slope, eps, residual_is_x=False):
        need_pre = (slope is not None or gamma is not None) and any(ctx.needs_input_grad[:3])
        pre = torch.empty((a.shape[0], x.shape[1]), dtype=torch.float32, device=x.device) if need_pre else None
        g = gamma.contiguous() if gamma is not None else None
        b = beta.contiguous() if beta is not None else None
        if residual_is_x:  # y = f(x) + x (EquivSetConv's two convolutions): the residual gradient rides on the backward epilogue
            residual = x
        ctx.res_is_x = residual_is_x
        r = _check_dense(residual, a.shape[0], "residual") if residual is not None else None
        y = hgconv_raw(a, a.t(), x, _epilogue(slope=slope, gamma=g, beta=b, eps=eps, residual=r, pre=pre))
        ctx.a, ctx.slope, ctx.eps = a, slope, eps
        ctx.has_ln, ctx.has_res = gamma is not None, residual is not None
        ctx.save_for_backward(pre, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        pre, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        dgamma = dbeta = None
        if pre is not None:
            dz, dgamma, dbeta = leaky_ln_bwd(pre, dy, gamma if ctx.has_ln else None, ctx.eps, ctx.slope)
        else:
            dz = dy
        # y = f(A (At x))  =>  dx = At^T (A^T dz) = A (At dz) evaluated with the roles swapped
        if ctx.res_is_x:
            # dx = A (At dz) + dy in the second propagation's epilogue instead of a separate elementwise add
            dx = hgconv_raw(ctx.a, ctx.a.t(), dz, _epilogue(residual=dy)) if ctx.needs_input_grad[0] else None
            return dx, dgamma, dbeta, None, None, None, None, None
        dx = _hgconv_transposed(ctx.a, dz) if ctx.needs_input_grad[0] else None
        return dx, dgamma, dbeta, (dy if ctx.has_res else None), None, None, None, None


def a_is_square(a) -> bool:
    return a.shape[0] == a.shape[1]


def _hgconv_transposed(a: DeviceCSR, dz: torch.Tensor) -> torch.Tensor:
    """Gradient of ``A (A^T x)`` w.r.t. ``x`` applied to ``dz``: ``A (A^T dz)`` (the operator is
    symmetric even when A is not)."""
    return hgconv_raw(a, a.t(), dz)


def hgconv(adj: DeviceCSR, x: torch.Tensor, slope: float | None = None, ln_weight: torch.Tensor | None = None,
           ln_bias: torch.Tensor | None = None, residual: torch.Tensor | None = None, eps: float = 1e-5) -> torch.Tensor:
    """``[LayerNorm](leaky_relu(adj @ (adj.t() @ x))) [+ residual]`` in two kernel launches.

    ``slope=None`` is HGCNConv's ``act=False`` branch; ``ln_weight/ln_bias`` fuse the ``lns[k]`` that
    wraps every HGCNConv call in the reference; ``residual`` fuses the ``+ res``."""
    if (ln_weight is None) != (ln_bias is None):
        raise ValueError("ln_weight and ln_bias must be given together")
    if _sharded(adj):
        return adj.hgconv(x, slope, ln_weight, ln_bias, residual, eps)
    if residual is x and residual is not None and a_is_square(adj):
        return _HGConv.apply(x, ln_weight, ln_bias, None, adj, slope, eps, True)
    return _HGConv.apply(x, ln_weight, ln_bias, residual, adj, slope, eps)


def lightgcn_propagate_raw(a: DeviceCSR, e0: torch.Tensor, n_layers: int, sum_readout: bool = False) -> torch.Tensor:
    e0 = _check_dense(e0, a.shape[1], "E0")
    n, d = e0.shape
    out = torch.empty_like(e0)
    if n_layers == 0:
        return e0.clone()
    layers = torch.empty((max(n_layers - 1, 1), n, d), dtype=torch.float32, device=e0.device) if n_layers > 1 else None
    ws, ws_bytes = _ws(a, d)
    with _Timed(n_layers):
        _lib.check(_lib.lib().hgr_lightgcn_forward_f32(C.byref(a.desc), e0.data_ptr(), _lib.ptr(layers), out.data_ptr(), n_layers, d,
                                                       int(sum_readout), ws, ws_bytes, _lib.stream_ptr()))
    return out


class _LightGCN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e0, a, n_layers, sum_readout):
        ctx.a, ctx.n_layers, ctx.sum_readout = a, n_layers, sum_readout
        return lightgcn_propagate_raw(a, e0, n_layers, sum_readout)

    @staticmethod
    def backward(ctx, dout):
        # out = c * sum_k A^k e0  =>  de0 = c * sum_k (A^T)^k dout : the same fused call on A^T
        return lightgcn_propagate_raw(ctx.a.t(), dout.contiguous(), ctx.n_layers, ctx.sum_readout), None, None, None


def lightgcn_propagate(adj: DeviceCSR, ego: torch.Tensor, n_layers: int, sum_readout: bool = False) -> torch.Tensor:
    """``mean_k (adj^k @ ego)`` for k = 0..n_layers (``sum_readout``: the plain sum) -- the body of
    ``LGCN_Encoder.forward`` in ``n_layers`` launches; E^L and the stacked [N, L+1, D] tensor of the
    reference are never materialised."""
    if _sharded(adj):
        return adj.lightgcn_propagate(ego, n_layers, sum_readout)
    return _LightGCN.apply(ego, adj, n_layers, sum_readout)


def scatter_mean_conv(inc, x: torch.Tensor) -> torch.Tensor:
    """Scatter form of the node -> hyperedge -> node message passing (model/layers/layers2/EquivSetConv2.py:88-93
    with identity MLPs): ``Xe = scatter_mean_E(X[V])``, ``Xv = scatter_mean_V(Xe[E])`` as two propagations over the
    row-normalised incidence pair (``graph.Incidence``).  Deterministic (no atomics); differentiable."""
    return spmm(inc.to_nodes, spmm(inc.to_edges, x))


class _WeightedToEdges(torch.autograd.Function):
    """``Xe = segment_mean_E(X[V] * att)`` (model/graph/HD2.py:629-633): the node -> hyperedge propagation with one attention weight
    per (vertex, hyperedge) pair folded into the CSR values (``att / |e|``).  Backward: ``dX`` through the transposed matrix with
    the same values; ``d att`` is a sampled dense-dense product ``<dXe[e], X[v]> / |e|`` over the pairs (torch gathers)."""

    @staticmethod
    def forward(ctx, x, att, inc, vertex, edges, pos_e, pos_n):
        te, tn = inc.to_edges, inc.to_nodes
        a = att.reshape(-1).to(torch.float32)
        inv_e = te.values[pos_e]  # 1 / |e| of every pair (the row-normalised incidence)
        vals = torch.zeros_like(te.values)
        vals[pos_e] = inv_e * a
        t_vals = torch.zeros_like(tn.values)
        t_vals[pos_n] = inv_e * a
        ctx.w = te.with_values(vals)
        # the transposed weighted matrix lives on to_nodes' pattern ([n_nodes, n_edges], canonical)
        ctx.wt = tn.with_values(t_vals)
        ctx.save_for_backward(x, inv_e, vertex, edges)
        ctx.att_shape = att.shape
        return spmm_raw(ctx.w, x)

    @staticmethod
    def backward(ctx, dxe):
        x, inv_e, vertex, edges = ctx.saved_tensors
        dxe = dxe.contiguous()
        dx = spmm_raw(ctx.wt, dxe) if ctx.needs_input_grad[0] else None
        datt = None
        if ctx.needs_input_grad[1]:
            datt = ((dxe[edges] * x[vertex]).sum(-1) * inv_e).reshape(ctx.att_shape)
        return dx, datt, None, None, None, None, None


def scatter_mean_conv_weighted(inc, x: torch.Tensor, att: torch.Tensor, vertex: torch.Tensor, edges: torch.Tensor, positions=None):
    """``scatter_mean_conv`` with a per-pair attention weight on the node -> hyperedge stage (HD2's ``EquivSetConv.forward``,
    model/graph/HD2.py:624-643).  ``positions`` = ``graph.pair_positions(inc, vertex, edges)`` when the caller has them cached."""
    from .graph import pair_positions

    pos_e, pos_n = positions if positions is not None else pair_positions(inc, vertex, edges)
    xe = _WeightedToEdges.apply(x, att, inc, vertex.to(x.device).long(), edges.to(x.device).long(), pos_e, pos_n)
    return spmm(inc.to_nodes, xe)


class _PatternMeanConv(torch.autograd.Function):
    """Scatter-mean message passing through a DENSE 0/1 incidence ``b [n, K]`` (HCCF_diffusion: the sign pattern of the learned
    ``E W``, model/graph/HCCF_diffusion.py:291-308,382-402): ``Xe = diag(1 / max(|e|, 1)) b^T X``, ``Xv = diag(1 / max(|v|, 1)) b Xe`` on
    the tall-and-skinny kernels (csrc/hyperedge.cu) instead of ``nonzero`` + two ``torch_scatter`` calls over ~n K / 2 pairs.
    ``b`` carries no gradient (the reference's ``nonzero(H > 0)`` cuts it)."""

    @staticmethod
    def forward(ctx, x, b):
        b = b.contiguous()
        inv_e = 1.0 / b.sum(0).clamp(min=1.0)
        inv_v = 1.0 / b.sum(1).clamp(min=1.0)
        xe = tall_skinny_tn(b, x.contiguous()) * inv_e[:, None]
        ctx.save_for_backward(b, inv_e, inv_v)
        return rows_times_small(b, None, xe.contiguous()) * inv_v[:, None]

    @staticmethod
    def backward(ctx, dy):
        b, inv_e, inv_v = ctx.saved_tensors
        g = (dy * inv_v[:, None]).contiguous()
        dxe = tall_skinny_tn(b, g) * inv_e[:, None]
        return rows_times_small(b, None, dxe.contiguous()), None


def pattern_mean_conv(b: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """See ``_PatternMeanConv``; ``b`` float32 ``[n, K]`` of zeros and ones with K in {32, 64, 128, 256}, ``x`` ``[n, 32|64|128]``."""
    if not hyperedge_supported(b, x):
        raise ValueError("pattern_mean_conv: need float32 CUDA b [n, 32|64|128|256] and x [n, 32|64|128], got %s and %s" % (tuple(b.shape), tuple(x.shape)))
    return _PatternMeanConv.apply(x, b)


def segment_mean_to_edges(inc, x: torch.Tensor) -> torch.Tensor:
    """``torch_scatter.scatter(X[V], E, dim=-2, reduce='mean')``"""
    return spmm(inc.to_edges, x)


def segment_mean_to_nodes(inc, xe: torch.Tensor) -> torch.Tensor:
    """``torch_scatter.scatter(Xe[E], V, dim=-2, reduce='mean', dim_size=N)``"""
    return spmm(inc.to_nodes, xe)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x, w, b = x.contiguous(), weight.contiguous(), bias.contiguous()
        y = torch.empty_like(x)
        _lib.check(_lib.lib().hgr_layer_norm_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), float(eps), x.shape[0], x.shape[1], y.data_ptr(),
                                                 _lib.stream_ptr()))
        ctx.save_for_backward(x, w)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx, dgamma, dbeta = leaky_ln_bwd(x, dy.contiguous(), w, ctx.eps, None)
        return dx, dgamma, dbeta, None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``torch.nn.functional.layer_norm(x, (D,), weight, bias, eps)`` for 2-D float32 CUDA rows of width 32 / 64 / 128 on the
    row-group kernels of csrc/rowwise.cu (same arithmetic as the propagation epilogue)."""
    return _LayerNorm.apply(x, weight, bias, eps)


# ------------------------------------------------------------------------------------------------
# dense learned-hyperedge propagation of HCCF (csrc/hyperedge.cu) -- HGNNLayer.forward, model/graph/HCCF.py:206-211
# ------------------------------------------------------------------------------------------------
def tall_skinny_tn(h: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """``h.T @ e`` for ``h [n, K]``, ``e [n, D]`` with n large: every SM reduces a slice of the rows (hgr_tall_skinny_tn_f32)."""
    n, k = h.shape
    d = e.shape[1]
    lib = _lib.lib()
    t = torch.empty((k, d), dtype=torch.float32, device=h.device)
    ws_bytes = int(lib.hgr_tall_skinny_workspace_bytes(n, k, d))
    ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.float32, device=h.device)
    _lib.check(lib.hgr_tall_skinny_tn_f32(h.data_ptr(), e.data_ptr(), n, k, d, t.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
    return t


def rows_times_small(a1: torch.Tensor, a2: torch.Tensor | None, b: torch.Tensor, bias: torch.Tensor | None = None,
                     relu: bool = False, gather=None) -> torch.Tensor:
    """``[relu](cat([a1, a2], 1) @ b [+ bias])`` for a small ``b`` kept in shared memory (hgr_rows_times_small_gather_f32);
    ``gather`` = ``(peer pointers, row offset, multicast pointer)``: the rows are also published into every rank's gathered table."""
    n, k1 = a1.shape
    k2 = 0 if a2 is None else a2.shape[1]
    y = torch.empty((n, b.shape[1]), dtype=torch.float32, device=a1.device)
    g = None if gather is None else C.byref(_lib.make_gather(*gather))
    _lib.check(_lib.lib().hgr_rows_times_small_gather_f32(a1.data_ptr(), k1, _lib.ptr(a2), k2, b.data_ptr(), b.shape[1], n, y.data_ptr(),
                                                          _lib.ptr(bias), 1 if relu else 0, g, _lib.stream_ptr()))
    return y


class _AddRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, publish):
        a, b = a.contiguous(), b.contiguous()

        def run(gather):
            out = torch.empty_like(a)
            g = None if gather is None else C.byref(_lib.make_gather(*gather))
            _lib.check(_lib.lib().hgr_add_rows_f32(a.data_ptr(), b.data_ptr(), a.shape[0], a.shape[1], out.data_ptr(), g, _lib.stream_ptr()))
            return out

        return publish.publish_by(a.shape[1], run) if publish is not None else run(None)

    @staticmethod
    def backward(ctx, dy):
        return dy, dy, None


def add_rows(a: torch.Tensor, b: torch.Tensor, publish=None) -> torch.Tensor:
    """``a + b`` for two ``[rows, D]`` float32 CUDA tables on ``hgr_add_rows_f32``; ``publish`` (a sharded ``dist.DistGraph``): the sum
    is also stored into every rank's gathered table by the same kernel, so the propagation that consumes it needs no exchange."""
    if publish is not None and not (getattr(publish, "fused", False) and a.shape[0] == publish.part.n_loc):
        publish = None
    if not (a.is_cuda and a.dtype == b.dtype == torch.float32 and a.dim() == 2 and a.shape == b.shape and a.shape[1] % 4 == 0):
        return a + b
    return _AddRows.apply(a, b, publish)


def hyperedge_supported(h: torch.Tensor, e: torch.Tensor) -> bool:
    return (h.is_cuda and e.is_cuda and h.dtype == e.dtype == torch.float32 and h.dim() == e.dim() == 2 and h.shape[0] == e.shape[0]
            and h.shape[1] in (32, 64, 128, 256) and e.shape[1] in (32, 64, 128))


class _Hyperedge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, e):
        h, e = h.contiguous(), e.contiguous()
        t = tall_skinny_tn(h, e)
        ctx.save_for_backward(h, e, t)
        return rows_times_small(h, None, t)

    @staticmethod
    def backward(ctx, dy):
        h, e, t = ctx.saved_tensors
        dy = dy.contiguous()
        dt = tall_skinny_tn(h, dy)
        de = rows_times_small(h, None, dt) if ctx.needs_input_grad[1] else None
        dh = None
        if ctx.needs_input_grad[0]:
            d, k = e.shape[1], h.shape[1]
            if (2 * d * k + 2 * 32 * 2 * d) * 4 <= 220 * 1024:  # [T^T ; dT^T] and two (32-row) tiles of [dY | E] fit shared memory
                dh = rows_times_small(dy, e, torch.cat([t.t(), dt.t()], 0).contiguous())  # dY T^T + E dT^T in one pass
            else:
                dh = rows_times_small(dy, None, t.t().contiguous()) + rows_times_small(e, None, dt.t().contiguous())
        return dh, de


def hyperedge(h: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """``h @ (h.T @ e)``: node -> learned hyperedge -> node through the dense incidence ``h`` (HGNNLayer.forward)."""
    if not hyperedge_supported(h, e):
        raise ValueError("hyperedge: need float32 CUDA h [n, 32|64|128|256] and e [n, 32|64|128], got %s and %s" % (tuple(h.shape), tuple(e.shape)))
    return _Hyperedge.apply(h, e)


class _TallTimesSmall(torch.autograd.Function):
    """``x @ w`` for a tall ``x [n, K]`` and a small ``w [K, N]`` (HCCF's ``hyper = E0 @ W``, model/graph/HCCF.py:178-179):
    forward ``rows_times_small``, backward ``dx = dy @ w.T`` (same kernel) and ``dw = x.T @ dy`` (the tall-skinny reduce)."""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.contiguous(), w.contiguous()
        ctx.save_for_backward(x, w)
        return rows_times_small(x, None, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = rows_times_small(dy, None, w.t().contiguous()) if ctx.needs_input_grad[0] else None
        dw = tall_skinny_tn(x, dy) if ctx.needs_input_grad[1] else None
        return dx, dw


def tall_times_small_supported(x: torch.Tensor, w: torch.Tensor) -> bool:
    return (x.is_cuda and w.is_cuda and x.dtype == w.dtype == torch.float32 and x.dim() == w.dim() == 2 and x.shape[1] == w.shape[0]
            and x.shape[1] in (32, 64, 128) and w.shape[1] in (32, 64, 128))


def tall_times_small(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``x @ w`` on libhgr for ``x [n, 32|64|128]`` and ``w`` of width 32 | 64 | 128; anything else raises."""
    if not tall_times_small_supported(x, w):
        raise ValueError("tall_times_small: unsupported shapes %s @ %s" % (tuple(x.shape), tuple(w.shape)))
    return _TallTimesSmall.apply(x, w)


class _Linear(torch.autograd.Function):
    """``[relu](F.linear(x, weight, bias))`` for tall inputs.  Forward: one pass of the rows x small-matrix kernel with the bias
    (and the ReLU) in its epilogue, instead of cuBLAS sgemm + bias kernel + ReLU kernel (0.37 + 0.40 + 0.11 ms for 1.5 M rows of
    64).  Backward: ``dx = dy W`` on the same rows x small-matrix kernel, the weight gradient ``dy.T @ x`` -- a contraction over ALL
    rows that cuBLAS runs as a few-block SIMT kernel (0.84 ms) -- on the tall-skinny reduce of csrc/hyperedge.cu.  No library GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, publish=None):
        x = x.contiguous()
        wt, bb = weight.t().contiguous(), None if bias is None else bias.contiguous()
        if publish is not None:  # sharded: the output rows go into every rank's gathered table from this kernel's epilogue
            y = publish.publish_by(wt.shape[1], lambda gather: rows_times_small(x, None, wt, bb, relu, gather))
        else:
            y = rows_times_small(x, None, wt, bb, relu)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.has_bias, ctx.relu = bias is not None, relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.relu:
            dy = dy * (y > 0)
        dx = rows_times_small(dy, None, weight.contiguous()) if ctx.needs_input_grad[0] else None  # dy [n, out] x weight [out, in]
        dw = tall_skinny_tn(dy, x) if ctx.needs_input_grad[1] else None
        db = dy.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, relu: bool = False, publish=None) -> torch.Tensor:
    """Drop-in for ``nn.Linear.forward`` (``relu=True``: followed by ``F.relu``) on ``[rows, in]`` float32 CUDA inputs:
    out in {32, 64, 128, 256}, in in {32, 64, 128} -- the widths of every Linear on the hot path (MLP / lin_in of the ED-HNN
    blocks).  Anything else raises: there is no library fallback behind it."""
    if not x.is_cuda:
        raise _lib.HgrError("linear: input must be a CUDA tensor (no CPU path)")
    if not (x.dim() == 2 and x.dtype == weight.dtype == torch.float32 and weight.shape[0] in (32, 64, 128, 256)
            and weight.shape[1] in (32, 64, 128) and x.shape[1] == weight.shape[1]):
        raise ValueError("linear: need float32 x [rows, 32|64|128] and weight [32|64|128|256, in], got %s and %s" % (tuple(x.shape), tuple(weight.shape)))
    if publish is not None and not (getattr(publish, "fused", False) and x.shape[0] == publish.part.n_loc):
        publish = None  # only a sharded graph with the fused exchange can carry the rows
    return _Linear.apply(x, weight, bias, relu, publish)
