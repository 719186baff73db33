"""Autograd wiring of the training path (BASELINE config 4): `torch.autograd.Function`s whose forward
AND backward are kernels of librdm_sm100.so (csrc/backward.cu). `rdmnet_b200.ops` routes through them whenever gradients
are enabled and an input requires one, so the module mirror (rdmnet_b200/modules.py) - and with it the reference's
unmodified experiments/{model,loss,trainval}.py on top of rdmnet_b200.dropin - trains through our kernels.

What the reference does here: plain PyTorch autograd over its ATen graphs (experiments/trainval.py:43-50,
geotransformer/engine/epoch_based_trainer.py:104 loss.backward()).

Coverage: KPConv (gather + weight contraction), Linear (+ fused activation), GroupNorm (+ residual + LeakyReLU), LayerNorm
(+ residual + ReLU), max-pool, nearest-upsample + concat, index_select, the score activations (csrc/backward.cu); the rotary
embedding, the multi-head attention core and the 100 unrolled Sinkhorn iterations (csrc/train.cu).
"""
import torch

from . import _lib as L


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors)


def _ws(n, dev):
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=dev)


def _raw_linear(x, w, bias, b_is_nk, act=0):
    """C = act(x @ (w^T if b_is_nk else w) + bias) through rdm_linear (no autograd). A [K,N]-layout weight gets the workspace that
    lets the library transpose it and stay on the tensor cores."""
    x = x.contiguous()
    w = w.contiguous()
    m, k = x.shape
    n = w.shape[0] if b_is_nk else w.shape[1]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    if b_is_nk:
        wsb = L.lib().rdm_linear_workspace(m, n, k) if m * n <= (1 << 20) else 0
    else:
        wsb = L.lib().rdm_linear_kn_workspace(m, n, k)
    ws = _ws(wsb, x.device) if wsb else None
    L.call("rdm_linear", L.ptr(x), k, L.ptr(w), w.shape[1], 1 if b_is_nk else 0, L.ptr(bias), L.ptr(out), n, m, n, k, act,
           L.ptr(ws), wsb, L.stream())
    return out


def transpose(x):
    x = x.contiguous()
    r, c = x.shape
    y = torch.empty((c, r), dtype=torch.float32, device=x.device)
    L.call("rdm_transpose", L.ptr(x), r, c, c, L.ptr(y), L.stream())
    return y


def _act_bwd(y, dy, act, slope=0.1):
    dx = torch.empty_like(dy)
    L.call("rdm_activation_bwd", L.ptr(y.contiguous()), L.ptr(dy.contiguous()), dy.numel(), act, slope, L.ptr(dx), L.stream())
    return dx


class Linear(torch.autograd.Function):
    """y = act(x @ W^T + b) (weight_is_kn=False, nn.Linear layout) or act(x @ W + b) (weight_is_kn=True, KPConv layout)."""

    @staticmethod
    def forward(ctx, x, weight, bias, weight_is_kn, act):
        y = _raw_linear(x, weight, bias, not weight_is_kn, act)
        ctx.save_for_backward(x, weight, y if act else None)
        ctx.kn, ctx.act, ctx.has_bias = weight_is_kn, act, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        x, w, dy = x.contiguous(), w.contiguous(), dy.contiguous()
        if ctx.act:
            dy = _act_bwd(y, dy, ctx.act)
        m, k = x.shape
        n = dy.shape[1]
        dev = dy.device
        dx = torch.empty((m, k), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros(n, dtype=torch.float32, device=dev) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        wsb = L.lib().rdm_linear_bwd_workspace(m, n, k)
        ws = _ws(wsb, dev)
        L.call("rdm_linear_bwd", L.ptr(x), L.ptr(w), 0 if ctx.kn else 1, L.ptr(dy), m, n, k, L.ptr(dx), L.ptr(dw), L.ptr(db), L.ptr(ws), wsb,
               L.stream())
        return dx, dw, db, None, None


class KPConvGather(torch.autograd.Function):
    """A[m, k*C + c] of kpconv.py:88-116 (gather half of KPConv.forward); gradient w.r.t. the support features only (points
    and kernel points are not trained; the neighbour count is a comparison)."""

    @staticmethod
    def forward(ctx, s_feats, q_points, s_points, idx, kernel_points, h_kernel_points, sigma, query_order):
        from . import ops
        m, h = idx.shape
        n, c = s_feats.shape
        out = torch.empty((m, 15 * c), dtype=torch.float32, device=s_feats.device)
        rowpos = torch.empty(max(n, 1), dtype=torch.uint8, device=s_feats.device)
        L.call("rdm_kpconv_gather", L.ptr(s_feats), L.ptr(q_points), L.ptr(s_points), L.ptr(idx), ops._idx_bytes(idx), L.ptr(kernel_points),
               h_kernel_points.data_ptr(), float(sigma), m, n, h, c, L.ptr(query_order), L.ptr(out), L.ptr(rowpos), L.stream())
        ctx.save_for_backward(s_feats, q_points, s_points, idx)
        ctx.hk, ctx.sigma = h_kernel_points, float(sigma)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import ops
        s_feats, q_points, s_points, idx = ctx.saved_tensors
        m, h = idx.shape
        n, c = s_feats.shape
        d_feats = torch.empty_like(s_feats)
        rowpos = torch.empty(max(n, 1), dtype=torch.uint8, device=s_feats.device)
        L.call("rdm_kpconv_gather_bwd", L.ptr(d_out.contiguous()), L.ptr(s_feats), L.ptr(q_points), L.ptr(s_points), L.ptr(idx),
               ops._idx_bytes(idx), ctx.hk.data_ptr(), ctx.sigma, m, n, h, c, L.ptr(rowpos), L.ptr(d_feats), L.stream())
        return d_feats, None, None, None, None, None, None, None


class GroupNorm(torch.autograd.Function):
    """y = act(GroupNorm(x) * gamma + beta (+ residual)) over the stacked (N, C/G) slabs (kpconv/modules.py:33-50)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, groups, act, slope, eps):
        n, c = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(2 * groups, dtype=torch.float64, device=x.device)
        L.call("rdm_groupnorm", L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(residual), L.ptr(y), n, c, groups, eps, act, slope, L.ptr(stats),
               L.stream())
        ctx.save_for_backward(x, gamma, y)
        ctx.cfg = (groups, act, slope, eps, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, y = ctx.saved_tensors
        groups, act, slope, eps, has_res = ctx.cfg
        n, c = x.shape
        dev = x.device
        dx, dz = torch.empty_like(x), torch.empty_like(x)
        dg, db = torch.empty(c, dtype=torch.float32, device=dev), torch.empty(c, dtype=torch.float32, device=dev)
        stats = torch.empty(2 * groups, dtype=torch.float64, device=dev)
        scr = torch.empty(2 * c, dtype=torch.float64, device=dev)
        L.call("rdm_groupnorm_bwd", L.ptr(x), L.ptr(y), L.ptr(dy.contiguous()), L.ptr(gamma), n, c, groups, eps, act, slope, L.ptr(stats),
               L.ptr(scr), L.ptr(dz), L.ptr(dx), L.ptr(dg), L.ptr(db), L.stream())
        return dx, dg, db, (dz if has_res else None), None, None, None, None


class LayerNorm(torch.autograd.Function):
    """y = act(LayerNorm(x (+ residual)) * gamma + beta), act 2 = ReLU."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps, act):
        n, c = x.shape
        y = torch.empty_like(x)
        L.call("rdm_layernorm", L.ptr(x), L.ptr(residual), L.ptr(gamma), L.ptr(beta), L.ptr(y), n, c, eps, act, L.stream())
        ctx.save_for_backward(x, residual, gamma, y)
        ctx.cfg = (eps, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, residual, gamma, y = ctx.saved_tensors
        eps, act = ctx.cfg
        n, c = x.shape
        dx = torch.empty_like(x)
        dg = torch.zeros(c, dtype=torch.float32, device=x.device)
        db = torch.zeros(c, dtype=torch.float32, device=x.device)
        L.call("rdm_layernorm_bwd", L.ptr(x), L.ptr(residual), L.ptr(y), L.ptr(dy.contiguous()), L.ptr(gamma), n, c, eps, act, L.ptr(dx),
               L.ptr(dg), L.ptr(db), L.stream())
        return dx, (dx if residual is not None else None), dg, db, None, None


class MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        from . import ops
        m, h = idx.shape
        n, c = x.shape
        out = torch.empty((m, c), dtype=torch.float32, device=x.device)
        L.call("rdm_maxpool", L.ptr(x), L.ptr(idx), ops._idx_bytes(idx), m, n, h, c, L.ptr(out), L.stream())
        ctx.save_for_backward(x, idx)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import ops
        x, idx = ctx.saved_tensors
        m, h = idx.shape
        n, c = x.shape
        dx = torch.zeros_like(x)
        L.call("rdm_maxpool_bwd", L.ptr(x), L.ptr(idx), ops._idx_bytes(idx), L.ptr(d_out.contiguous()), m, n, h, c, L.ptr(dx), L.stream())
        return dx, None


class UpsampleConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, skip):
        from . import ops
        m = idx.shape[0]
        n, c1 = x.shape
        c2 = skip.shape[1] if skip is not None else 0
        out = torch.empty((m, c1 + c2), dtype=torch.float32, device=x.device)
        L.call("rdm_upsample_concat", L.ptr(x), idx.data_ptr(), ops._idx_bytes(idx), idx.stride(0), L.ptr(skip), m, n, c1, c2, L.ptr(out),
               L.stream())
        ctx.save_for_backward(idx)
        ctx.shape = (n, c1, c2)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from . import ops
        (idx,) = ctx.saved_tensors
        n, c1, c2 = ctx.shape
        m = idx.shape[0]
        dx = torch.zeros((n, c1), dtype=torch.float32, device=d_out.device)
        dskip = torch.empty((m, c2), dtype=torch.float32, device=d_out.device) if c2 else None
        L.call("rdm_upsample_concat_bwd", L.ptr(d_out.contiguous()), idx.data_ptr(), ops._idx_bytes(idx), idx.stride(0), m, n, c1, c2,
               L.ptr(dx), L.ptr(dskip), L.stream())
        return dx, None, dskip


class Activation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, slope):
        y = torch.empty_like(x)
        L.call("rdm_activation", L.ptr(x), L.ptr(y), x.numel(), act, slope, L.stream())
        ctx.save_for_backward(y)
        ctx.cfg = (act, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return _act_bwd(y, dy, ctx.cfg[0], ctx.cfg[1]), None, None


class IndexSelectRows(torch.autograd.Function):
    """index_select along dim 0 on a contiguous (rows, ...) float32 table; backward = rdm_scatter_add_rows."""

    @staticmethod
    def forward(ctx, data, flat_index):
        rows = data.shape[0]
        words = int(data.numel() // max(rows, 1))
        out = torch.empty((flat_index.shape[0],) + tuple(data.shape[1:]), dtype=data.dtype, device=data.device)
        if flat_index.shape[0] > 0 and words > 0:
            L.call("rdm_index_select", L.ptr(data), rows, words, L.ptr(flat_index), flat_index.element_size(), flat_index.shape[0],
                   L.ptr(out), None, L.stream())
        ctx.save_for_backward(flat_index)
        ctx.shape = tuple(data.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (flat_index,) = ctx.saved_tensors
        d = torch.zeros(ctx.shape, dtype=torch.float32, device=d_out.device)
        words = int(d.numel() // max(ctx.shape[0], 1))
        if flat_index.shape[0] > 0 and words > 0:
            L.call("rdm_scatter_add_rows", L.ptr(d_out.contiguous()), L.ptr(flat_index), flat_index.element_size(), flat_index.shape[0],
                   words, ctx.shape[0], L.ptr(d), L.stream())
        return d, None


class Rope(torch.autograd.Function):
    """RotaryPositionalEmbedding.forward (rdmnet/thdroformer/thdroformer.py:56-85): x (N,C), emb (N,C/2)."""

    @staticmethod
    def forward(ctx, x, emb):
        x, emb = x.contiguous(), emb.contiguous()
        n, c = x.shape
        y = torch.empty((n, c), dtype=torch.float32, device=x.device)
        L.call("rdm_rope", L.ptr(x), c, L.ptr(emb), emb.shape[1], L.ptr(y), c, n, c, L.stream())
        ctx.save_for_backward(x, emb)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, emb = ctx.saved_tensors
        n, c = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        demb = torch.empty_like(emb) if ctx.needs_input_grad[1] else None
        L.call("rdm_rope_bwd", L.ptr(x), c, L.ptr(emb), emb.shape[1], L.ptr(dy), c, n, c, L.ptr(dx), L.ptr(demb), L.stream())
        return dx, demb


class Attention(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v per head (thdroformer.py:20-40 with k = None; vanilla_transformer.py:54-66)."""

    @staticmethod
    def forward(ctx, q, k, v, heads):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        nq, c = q.shape
        nk = k.shape[0]
        o = torch.empty((nq, c), dtype=torch.float32, device=q.device)
        L.call("rdm_attention", L.ptr(q), c, L.ptr(k), c, L.ptr(v), c, L.ptr(o), c, nq, nk, heads, c // heads, L.stream())
        ctx.save_for_backward(q, k, v, o)
        ctx.heads = heads
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o = ctx.saved_tensors
        heads = ctx.heads
        nq, c = q.shape
        nk = k.shape[0]
        do = do.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        wsb = L.lib().rdm_attention_bwd_workspace(nq, heads)
        ws = _ws(wsb, q.device)
        L.call("rdm_attention_bwd", L.ptr(q), c, L.ptr(k), c, L.ptr(v), c, L.ptr(o), c, L.ptr(do), c, nq, nk, heads, c // heads, L.ptr(ws),
               wsb, L.ptr(dq), L.ptr(dk), L.ptr(dv), c, L.stream())
        return dq, dk, dv, None


class Sinkhorn(torch.autograd.Function):
    """LearnableLogOptimalTransport.forward (geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66): scores (P,R,C), masks
    (P,R) / (P,C) uint8 (1 = live), alpha 0-d -> (P,R+1,C+1). Backward: rdm_sinkhorn_bwd (gradient on masked entries ignored)."""

    @staticmethod
    def forward(ctx, scores, row_masks, col_masks, alpha, num_iterations, inf):
        scores = scores.contiguous()
        p, r, c = scores.shape
        out = torch.empty((p, r + 1, c + 1), dtype=torch.float32, device=scores.device)
        a = alpha.detach().reshape(1).contiguous()
        L.call("rdm_sinkhorn", L.ptr(scores), p, r, c, L.ptr(row_masks), L.ptr(col_masks), None, None, L.ptr(a), int(num_iterations),
               float(inf), L.ptr(out), L.stream())
        ctx.save_for_backward(scores, row_masks, col_masks, a)
        ctx.cfg = (int(num_iterations), float(inf), alpha.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        scores, rm, cm, a = ctx.saved_tensors
        iters, inf, ashape = ctx.cfg
        p, r, c = scores.shape
        d_scores = torch.empty_like(scores)
        part = torch.zeros(max(p, 1), dtype=torch.float32, device=scores.device)
        wsb = L.lib().rdm_sinkhorn_bwd_workspace(p, r, c, iters)
        ws = _ws(wsb, scores.device)
        L.call("rdm_sinkhorn_bwd", L.ptr(scores), p, r, c, L.ptr(rm), L.ptr(cm), L.ptr(a), iters, inf, L.ptr(d_out.contiguous()), L.ptr(ws),
               wsb, L.ptr(d_scores), L.ptr(part), L.stream())
        d_alpha = torch.zeros(1, dtype=torch.float32, device=scores.device)
        if p > 0:
            L.call("rdm_colsum", L.ptr(part), p, 1, 1, L.ptr(d_alpha), L.stream())
        return d_scores, None, None, d_alpha.reshape(ashape), None, None
