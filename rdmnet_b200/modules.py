"""Host-side mirror of the reference's NN building blocks for the hot path: same class names, constructor
signatures, forward signatures and state_dict keys as

  geotransformer.modules.kpconv.{KPConv, GroupNorm, UnaryBlock, LastUnaryBlock, ConvBlock, ResidualBlock}
  geotransformer.modules.transformer.{vanilla_transformer.TransformerLayer, output_layer.AttentionOutput}
  rdmnet.thdroformer.ThDRoFormer, rdmnet.vote.{Vote_layer, NMS}
  geotransformer.modules.sinkhorn.LearnableLogOptimalTransport
  geotransformer.modules.geotransformer.{SuperPointMatching, LocalGlobalRegistration}

so that a reference checkpoint loads with strict=True. The modules only own parameters; all arithmetic is done by
the CUDA kernels of librdm_sm100.so through rdmnet_b200.ops (inference path; autograd is not wired in round 1).
"""
import math

import torch
import torch.nn as nn

from . import ops


def default_kernel_points(radius, num_kpoints=15):
    """Deterministic 'center' disposition for fresh initialisation: one point at the origin and the others on a
    Fibonacci sphere of radius 0.66*radius. (The reference optimises a disposition offline and applies a random
    rotation - geotransformer/modules/kpconv/kernel_points.py:389-455 - which is init-time only: at inference the
    buffer is overwritten by the checkpoint.)"""
    pts = torch.zeros(num_kpoints, 3)
    n = num_kpoints - 1
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(n):
        z = 1 - 2 * (i + 0.5) / n
        r = math.sqrt(max(0.0, 1 - z * z))
        pts[i + 1] = torch.tensor([math.cos(golden * i) * r, math.sin(golden * i) * r, z])
    return pts * 0.66 * radius


_WEIGHTS_EPOCH = [0]


class _Module(nn.Module):
    """nn.Module whose device moves / dtype casts / state-dict loads / train-eval switches bump a process-wide epoch.
    Derived data (transposed weight copies, packed layer blobs, C descriptor structs) are cached against
    (mode, epoch, fingerprint): the epoch catches structural changes cheaply, the fingerprint - (storage pointer, in-place
    version) of every parameter and buffer of the module - catches what the epoch cannot see: `p.copy_()`, `p.data = ...`,
    an EMA swap, pruning, `load_state_dict` called directly on a plain nn.Linear child, an optimizer step. The tensor
    list behind the fingerprint is itself cached per epoch, so a check costs two attribute reads per tensor (~50 us for
    the whole 497-tensor model, once per forward: RDMNet.forward_head shares it with the runners)."""

    def _apply(self, fn, *args, **kwargs):
        _WEIGHTS_EPOCH[0] += 1
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        _WEIGHTS_EPOCH[0] += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def train(self, mode=True):
        _WEIGHTS_EPOCH[0] += 1
        return super().train(mode)


def invalidate_caches():
    """Drops every cached derived copy / descriptor (they are rebuilt on the next forward). The (pointer, version)
    fingerprint makes this unnecessary for ordinary in-place edits; it remains for exotic cases (storage edited through a
    raw pointer by foreign code)."""
    _WEIGHTS_EPOCH[0] += 1


def cache_key(module, tensors=None):
    """Cache key for data derived from `module`'s parameters (see _Module)."""
    if tensors is None:
        cached = module.__dict__.get("_fp_tensors")
        if cached is None or cached[0] != _WEIGHTS_EPOCH[0]:
            cached = (_WEIGHTS_EPOCH[0], list(module.parameters()) + list(module.buffers()))
            module.__dict__["_fp_tensors"] = cached
        tensors = cached[1]
    return (module.training, _WEIGHTS_EPOCH[0], tuple([(t.data_ptr(), t._version) for t in tensors]))


class KPConv(_Module):
    """geotransformer/modules/kpconv/kpconv.py:10-122."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, bias=False, dimension=3, inf=1e6, eps=1e-9):
        super().__init__()
        if kernel_size != 15 or dimension != 3:
            raise ValueError("rdmnet_b200.KPConv supports the 15-point 3-D kernel of the reference configuration")
        self.kernel_size, self.in_channels, self.out_channels = kernel_size, in_channels, out_channels
        self.radius, self.sigma, self.dimension, self.inf, self.eps = radius, sigma, dimension, inf, eps
        self.weights = nn.Parameter(torch.zeros(kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self.register_buffer("kernel_points", default_kernel_points(radius, kernel_size))

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weights)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, s_feats, q_points, s_points, neighbor_indices):
        return ops.kpconv(s_feats, q_points, s_points, neighbor_indices, self.weights, self.kernel_points, self.sigma,
                          self.bias)


class GroupNorm(_Module):
    """geotransformer/modules/kpconv/modules.py:33-50 (statistics over the whole stacked (N, C/G) slab)."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups, self.num_channels = num_groups, num_channels
        self.norm = nn.GroupNorm(num_groups, num_channels)  # parameter holder: keys norm.weight / norm.bias

    def forward(self, x, residual=None, act=0):
        return ops.group_norm(x, self.norm.weight, self.norm.bias, self.num_groups, residual, act, 0.1, self.norm.eps)


class UnaryBlock(_Module):
    """kpconv/modules.py:53-83."""

    def __init__(self, in_channels, out_channels, group_norm, has_relu=True, bias=True, layer_norm=False):
        super().__init__()
        if layer_norm:
            raise ValueError("layer_norm=True is not used by RDMNet and is not implemented")
        self.in_channels, self.out_channels, self.group_norm = in_channels, out_channels, group_norm
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)
        self.leaky_relu = nn.LeakyReLU(0.1) if has_relu else None

    def forward(self, x, residual=None, act=None):
        x = ops.linear(x, self.mlp.weight, self.mlp.bias)
        if act is None:
            act = 1 if self.leaky_relu is not None else 0
        return self.norm(x, residual, act)


class LastUnaryBlock(_Module):
    """kpconv/modules.py:86-101."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x):
        return ops.linear(x, self.mlp.weight, self.mlp.bias)


class ConvBlock(_Module):
    """kpconv/modules.py:104-147."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, group_norm, negative_slope=0.1, bias=True,
                 layer_norm=False):
        super().__init__()
        if layer_norm or negative_slope != 0.1:
            raise ValueError("only GroupNorm + LeakyReLU(0.1) (the RDMNet configuration) is implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.KPConv = KPConv(in_channels, out_channels, kernel_size, radius, sigma, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)
        self.leaky_relu = nn.LeakyReLU(negative_slope=negative_slope)

    def forward(self, s_feats, q_points, s_points, neighbor_indices):
        x = self.KPConv(s_feats, q_points, s_points, neighbor_indices)
        return self.norm(x, None, 1)


class ResidualBlock(_Module):
    """kpconv/modules.py:150-225."""

    def __init__(self, in_channels, out_channels, kernel_size, radius, sigma, group_norm, strided=False, bias=True,
                 layer_norm=False):
        super().__init__()
        if layer_norm:
            raise ValueError("layer_norm=True is not used by RDMNet and is not implemented")
        self.in_channels, self.out_channels, self.strided = in_channels, out_channels, strided
        mid = out_channels // 4
        self.unary1 = UnaryBlock(in_channels, mid, group_norm, bias=bias) if in_channels != mid else nn.Identity()
        self.KPConv = KPConv(mid, mid, kernel_size, radius, sigma, bias=bias)
        self.norm_conv = GroupNorm(group_norm, mid)
        self.unary2 = UnaryBlock(mid, out_channels, group_norm, has_relu=False, bias=bias)
        if in_channels != out_channels:
            self.unary_shortcut = UnaryBlock(in_channels, out_channels, group_norm, has_relu=False, bias=bias)
        else:
            self.unary_shortcut = nn.Identity()
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, s_feats, q_points, s_points, neighbor_indices):
        x = self.unary1(s_feats)
        x = self.KPConv(x, q_points, s_points, neighbor_indices)
        x = self.norm_conv(x, None, 1)
        shortcut = ops.maxpool(s_feats, neighbor_indices) if self.strided else s_feats
        shortcut = self.unary_shortcut(shortcut)
        # unary2 = Linear + GroupNorm, fused with "+ shortcut" and the final LeakyReLU (:222-224)
        return self.unary2(x, residual=shortcut, act=1)


# ------------------------------------------------------------------------------------------------- transformer
class AttentionOutput(_Module):
    """geotransformer/modules/transformer/output_layer.py:6-21 (dropout=None)."""

    def __init__(self, d_model, dropout=None, activation_fn="ReLU"):
        super().__init__()
        if dropout is not None or activation_fn != "ReLU":
            raise ValueError("only dropout=None / ReLU (the RDMNet configuration) is implemented")
        self.expand = nn.Linear(d_model, d_model * 2)
        self.squeeze = nn.Linear(d_model * 2, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x):
        h = ops.linear(x, self.expand.weight, self.expand.bias, act=2)
        h = ops.linear(h, self.squeeze.weight, self.squeeze.bias)
        return ops.layer_norm(h, self.norm.weight, self.norm.bias, residual=x, eps=self.norm.eps)


class _PosEncoderBuffers(_Module):
    """Holds the (unused) `div_term` buffer of RotaryPositionalEmbedding so that checkpoints load strictly
    (rdmnet/thdroformer/thdroformer.py:43-54)."""

    def __init__(self, d_model, num_heads):
        super().__init__()
        d = d_model // num_heads
        div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
        self.register_buffer("div_term", div.repeat_interleave(2).view(1, 1, 1, -1))


class MultiHeadAttention(_Module):
    """vanilla_transformer.py:15-70 (no masks / factors: RDMNet passes none) and, with `rotary=True`,
    RPEMultiHeadAttention (thdroformer.py:88-139, k=None)."""

    def __init__(self, d_model, num_heads, dropout=None, rotary=False):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError("`d_model` ({}) must be a multiple of `num_heads` ({}).".format(d_model, num_heads))
        self.d_model, self.num_heads, self.d_model_per_head = d_model, num_heads, d_model // num_heads
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        if rotary:
            self.pos_encoder = _PosEncoderBuffers(d_model, num_heads)
        self.rotary = rotary

    def forward(self, input_q, input_k, input_v, embed_q=None, embed_k=None):
        q = ops.linear(input_q, self.proj_q.weight, self.proj_q.bias)
        k = ops.linear(input_k, self.proj_k.weight, self.proj_k.bias)
        v = ops.linear(input_v, self.proj_v.weight, self.proj_v.bias)
        if self.rotary:
            q = ops.rope(q, embed_q)
            k = ops.rope(k, embed_k)
        return ops.attention(q, k, v, self.num_heads)


class AttentionLayer(_Module):
    """vanilla_transformer.py:73-102 / RPEAttentionLayer thdroformer.py:141-172."""

    def __init__(self, d_model, num_heads, dropout=None, rotary=False):
        super().__init__()
        self.attention = MultiHeadAttention(d_model, num_heads, dropout=dropout, rotary=rotary)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, input_states, memory_states, embed_q=None, embed_k=None):
        h = self.attention(input_states, memory_states, memory_states, embed_q, embed_k)
        h = ops.linear(h, self.linear.weight, self.linear.bias)
        return ops.layer_norm(h, self.norm.weight, self.norm.bias, residual=input_states, eps=self.norm.eps)


class TransformerLayer(_Module):
    """vanilla_transformer.py:105-129 / RPETransformerLayer thdroformer.py:175-202. 2-D (N,C) states."""

    def __init__(self, d_model, num_heads, dropout=None, activation_fn="ReLU", rotary=False):
        super().__init__()
        self.attention = AttentionLayer(d_model, num_heads, dropout=dropout, rotary=rotary)
        self.output = AttentionOutput(d_model, dropout=dropout, activation_fn=activation_fn)
        self._blob, self._blob_key = None, None

    def forward(self, input_states, memory_states, embed_q=None, embed_k=None):
        return self.output(self.attention(input_states, memory_states, embed_q, embed_k))

    # ---- fused path (librdm_sm100: rdm_tf_project / rdm_tf_attend), d_model 128, 4 heads, FFN 256
    def fusable(self):
        a = self.attention.attention
        return (a.d_model == 128 and a.num_heads == 4 and self.output.expand.out_features == 256
                and self.attention.norm.eps == 1e-5 and self.output.norm.eps == 1e-5)

    def fused_blob(self):
        """The layer's parameters packed for the fused kernels (weights transposed to [in][out]); cached until a
        parameter changes (storage pointer or in-place version)."""
        a, al, o = self.attention.attention, self.attention, self.output
        params = [a.proj_q.weight, a.proj_k.weight, a.proj_v.weight, al.linear.weight, o.expand.weight, o.squeeze.weight,
                  a.proj_q.bias, a.proj_k.bias, a.proj_v.bias, al.linear.bias, o.expand.bias, o.squeeze.bias,
                  al.norm.weight, al.norm.bias, o.norm.weight, o.norm.bias]
        key = cache_key(self, params)
        if self._blob_key != key:
            with torch.no_grad():
                parts = [p.detach().t().contiguous().reshape(-1) for p in params[:6]] + [p.detach().reshape(-1) for p in params[6:]]
                blob = torch.cat(parts).float().contiguous()
            if blob.numel() != ops.L.lib().rdm_tf_layer_blob_floats():
                raise RuntimeError("transformer layer blob layout mismatch")
            self._blob, self._blob_key = blob, key
        return self._blob


class RPEConditionalTransformer(_Module):
    """thdroformer.py:204-251: alternating self (rotary) / cross layers; cross attention is sequential
    (feats1 attends to the already-updated feats0, :244-245)."""

    def __init__(self, blocks, d_model, num_heads, dropout=None, activation_fn="ReLU", return_attention_scores=False,
                 parallel=False, k=None):
        super().__init__()
        if k is not None or parallel or return_attention_scores:
            raise ValueError("k / parallel / return_attention_scores are not used by RDMNet and are not implemented")
        self.blocks = blocks
        self.layers = nn.ModuleList(
            [TransformerLayer(d_model, num_heads, dropout, activation_fn, rotary=(b == "self")) for b in blocks])

    def forward(self, feats0, feats1, embeddings0, embeddings1, masks0=None, masks1=None):
        if (all(layer.fusable() for layer in self.layers) and feats0.shape[0] > 0 and feats1.shape[0] > 0
                and not ops.AG.needs_grad(feats0, feats1, *self.parameters())):
            return self._forward_fused(feats0.contiguous(), feats1.contiguous(), embeddings0.contiguous(),
                                       embeddings1.contiguous())
        for layer, block in zip(self.layers, self.blocks):
            if block == "self":
                feats0 = layer(feats0, feats0, embeddings0, embeddings0)
                feats1 = layer(feats1, feats1, embeddings1, embeddings1)
            else:
                feats0 = layer(feats0, feats1)
                feats1 = layer(feats1, feats0)
        return feats0, feats1

    def _forward_fused(self, f0, f1, e0, e1):
        """Two launches per (layer, independent problem set) instead of ~25: see csrc/transformer.cu."""
        n0, n1, dev = f0.shape[0], f1.shape[0], f0.device
        qv0 = torch.empty((2, n0, 128), dtype=torch.float32, device=dev)
        qv1 = torch.empty((2, n1, 128), dtype=torch.float32, device=dev)
        q0, v0, q1, v1 = qv0[0], qv0[1], qv1[0], qv1[1]
        # projected keys are kept channel-major (128, n padded to 4): coalesced, conflict-free K tiles in rdm_tf_attend
        k0 = torch.empty((128, (n0 + 3) // 4 * 4), dtype=torch.float32, device=dev)  # pad columns are masked
        k1 = torch.empty((128, (n1 + 3) // 4 * 4), dtype=torch.float32, device=dev)
        for layer, block in zip(self.layers, self.blocks):
            blob = layer.fused_blob()
            if block == "self":
                ops.tf_project(blob, [(f0, "q", e0, q0), (f0, "k", e0, k0), (f0, "v", None, v0),
                                      (f1, "q", e1, q1), (f1, "k", e1, k1), (f1, "v", None, v1)])
                o0, o1 = torch.empty_like(q0), torch.empty_like(q1)
                ops.tf_attend(blob, [(q0, k0, v0, f0, o0), (q1, k1, v1, f1, o1)])
                f0, f1 = o0, o1
            else:
                ops.tf_project(blob, [(f0, "q", None, q0), (f1, "k", None, k1), (f1, "v", None, v1)])
                o0 = torch.empty_like(q0)
                ops.tf_attend(blob, [(q0, k1, v1, f0, o0)])
                f0 = o0
                ops.tf_project(blob, [(f1, "q", None, q1), (f0, "k", None, k0), (f0, "v", None, v0)])
                o1 = torch.empty_like(q1)
                ops.tf_attend(blob, [(q1, k0, v0, f1, o1)])
                f1 = o1
        return f0, f1


class posEmbedding(_Module):
    """thdroformer.py:253-263."""

    def __init__(self, hidden_dim, reduction_a="max"):
        super().__init__()
        self.proj = nn.Linear(3, hidden_dim // 2)

    def forward(self, points):
        return ops.linear(points, self.proj.weight, self.proj.bias)


class ThDRoFormer(_Module):
    """rdmnet/thdroformer/thdroformer.py:266-347. Accepts (1,N,3)/(1,N,C) like the reference (or 2-D tensors) and
    returns tensors of the same rank."""

    def __init__(self, input_dim, output_dim, hidden_dim, num_heads, num_layers, k=None, dropout=None,
                 activation_fn="ReLU", reduction_a="max"):
        super().__init__()
        self.embedding = posEmbedding(hidden_dim, reduction_a=reduction_a)
        self.in_proj = nn.Linear(input_dim, hidden_dim)
        self.transformer = RPEConditionalTransformer(["self", "cross"] * num_layers, hidden_dim, num_heads,
                                                     dropout=dropout, activation_fn=activation_fn, k=k)
        self.out_proj = nn.Linear(hidden_dim, output_dim)

    def forward(self, ref_points, src_points, ref_feats, src_feats, ref_masks=None, src_masks=None):
        batched = ref_feats.ndim == 3
        if batched:
            if ref_feats.shape[0] != 1:
                raise RuntimeError("ThDRoFormer processes one pair per call (as the reference: thdroformer.py:76)")
            ref_points, src_points, ref_feats, src_feats = ref_points[0], src_points[0], ref_feats[0], src_feats[0]
        tr = self.transformer
        training_grad = ops.AG.needs_grad(ref_feats, src_feats, *self.parameters())  # per-operator (differentiable) path
        if all(layer.fusable() for layer in tr.layers) and ref_feats.shape[0] > 0 and src_feats.shape[0] > 0 and not training_grad:
            f0, f1 = self._forward_runner(ref_points.contiguous(), src_points.contiguous(), ref_feats, src_feats)
            return (f0[None], f1[None]) if batched else (f0, f1)
        e0, e1 = self.embedding(ref_points.contiguous()), self.embedding(src_points.contiguous())
        f0 = ops.linear(ref_feats, self.in_proj.weight, self.in_proj.bias)
        f1 = ops.linear(src_feats, self.in_proj.weight, self.in_proj.bias)
        f0, f1 = self.transformer(f0, f1, e0, e1)
        f0 = ops.linear(f0, self.out_proj.weight, self.out_proj.bias)
        f1 = ops.linear(f1, self.out_proj.weight, self.out_proj.bias)
        return (f0[None], f1[None]) if batched else (f0, f1)


    def runner_desc(self):
        """The rdm_thdroformer_desc of this module (cached; see _Module)."""
        L = ops.L
        tr = self.transformer
        key = cache_key(self)
        if getattr(self, "_desc_key", None) != key:
            blobs = [layer.fused_blob() for layer in tr.layers]
            from .model import _presplit, _presplit_reset
            ps_keep = []
            _presplit_reset(self)
            _presplit(self, self.in_proj.weight.detach(), ps_keep)
            _presplit(self, self.out_proj.weight.detach(), ps_keep)
            d = L.ThdroformerDesc()
            d.emb_w, d.emb_b = self.embedding.proj.weight.data_ptr(), self.embedding.proj.bias.data_ptr()
            d.in_w, d.in_b = self.in_proj.weight.data_ptr(), self.in_proj.bias.data_ptr()
            d.out_w, d.out_b = self.out_proj.weight.data_ptr(), self.out_proj.bias.data_ptr()
            for i, (b, kind) in enumerate(zip(blobs, tr.blocks)):
                d.layer_blobs[i], d.is_self[i] = b.data_ptr(), 1 if kind == "self" else 0
            d.num_layers, d.c_in, d.c_out = len(blobs), self.in_proj.in_features, self.out_proj.out_features
            self._desc, self._desc_key, self._desc_keep = d, key, ps_keep
        return self._desc

    def _forward_runner(self, rp, sp, rf, sf):
        """rdm_thdroformer_forward: embedding, in_proj, all fused layers and out_proj in one host call."""
        import ctypes
        L = ops.L
        d = self.runner_desc()
        if rf.stride(1) != 1 or sf.stride(1) != 1 or rf.shape[1] != d.c_in:
            raise RuntimeError("ThDRoFormer: feature tensors must be (N, input_dim) with contiguous channels")
        n0, n1, dev = rf.shape[0], sf.shape[0], rf.device
        o0 = torch.empty((n0, d.c_out), dtype=torch.float32, device=dev)
        o1 = torch.empty((n1, d.c_out), dtype=torch.float32, device=dev)
        wsb = L.lib().rdm_thdroformer_workspace(n0, n1, d.c_out)
        ws = torch.empty(int(wsb), dtype=torch.uint8, device=dev)
        L.call("rdm_thdroformer_forward", ctypes.byref(d), L.ptr(rp), n0, L.ptr(sp), n1, rf.data_ptr(), rf.stride(0),
               sf.data_ptr(), sf.stride(0), L.ptr(o0), L.ptr(o1), L.ptr(ws), int(wsb), L.stream())
        return o0, o1


# ------------------------------------------------------------------------------------------------- vote / NMS
class Vote_layer(_Module):
    """rdmnet/vote/vote.py:43-117. `cfgs` needs MLPS, MAX_TRANSLATE_RANGE, input_feats_dim."""

    def __init__(self, cfgs, r):
        super().__init__()
        pre = cfgs.input_feats_dim
        layers = []
        for width in cfgs.MLPS:
            layers.extend([nn.Linear(pre, width), nn.LayerNorm(width), nn.ReLU()])
            pre = width
        self.mlp_modules = nn.Sequential(*layers) if layers else None
        self.ctr_reg = nn.Linear(pre, 3 + cfgs.input_feats_dim)
        rng = cfgs.MAX_TRANSLATE_RANGE
        self.max_offset_limit = torch.tensor(rng).float() / r if rng is not None else None
        self.out_proj = nn.Sequential(nn.LayerNorm(cfgs.input_feats_dim))

    def forward(self, xyz, features, aug_rotation=None):
        if xyz.ndim == 3:
            xyz, features = xyz[0], features[0]
        h = features
        if self.mlp_modules is not None:
            mods = list(self.mlp_modules)
            for i in range(0, len(mods), 3):
                h = ops.linear(h, mods[i].weight, mods[i].bias)
                h = ops.layer_norm(h, mods[i + 1].weight, mods[i + 1].bias, relu=True, eps=mods[i + 1].eps)
        off = ops.linear(h, self.ctr_reg.weight, self.ctr_reg.bias)
        ctr, feat_off = off[:, :3], off[:, 3:]
        if self.max_offset_limit is not None:
            lim = self.max_offset_limit.to(xyz.device)
            ctr = torch.minimum(torch.maximum(ctr, -lim), lim)  # the two torch.where clamps of vote.py:105-107
        vote_xyz = xyz + ctr
        ln = self.out_proj[0]
        new_features = ops.layer_norm(features, ln.weight, ln.bias, residual=feat_off.contiguous(), eps=ln.eps)
        return vote_xyz, new_features


class NMS(_Module):
    """rdmnet/vote/vote.py:6-40: radius search on the shifted nodes + greedy selection, both on the device."""

    def __init__(self, cfgs, neighbor_limits):
        super().__init__()
        self.NMS_radius = cfgs.NMS_radius
        self.neighbor_limits = int(neighbor_limits[-1])

    @torch.no_grad()
    def forward(self, nodes_dict, length_dict=None, overlap_score=None, features=None, split=None):
        nodes = nodes_dict.contiguous()
        if length_dict is None:
            length_dict = torch.tensor([nodes.shape[0]], dtype=torch.int64, device=nodes.device)
        lengths = length_dict.to(device=nodes.device, dtype=torch.int64)
        idx, _ = ops.radius_search_raw(nodes, nodes, lengths, lengths, self.NMS_radius, self.neighbor_limits,
                                       index_dtype=torch.int32)
        return ops.nms(idx, split)


# ------------------------------------------------------------------------------------------------- matching
class LearnableLogOptimalTransport(_Module):
    """geotransformer/modules/sinkhorn/learnable_sinkhorn.py:5-70."""

    def __init__(self, num_iterations, inf=1e12):
        super().__init__()
        self.num_iterations = num_iterations
        self.register_parameter("alpha", nn.Parameter(torch.tensor(1.0)))
        self.inf = inf

    def forward(self, scores, row_masks=None, col_masks=None):
        b, m, n = scores.shape
        if row_masks is None:
            row_masks = torch.ones((b, m), dtype=torch.bool, device=scores.device)
        if col_masks is None:
            col_masks = torch.ones((b, n), dtype=torch.bool, device=scores.device)
        return ops.sinkhorn(scores, row_masks, col_masks, self.alpha, self.num_iterations, self.inf)

    def __repr__(self):
        return self.__class__.__name__ + "(num_iterations={})".format(self.num_iterations)


class SuperPointMatching(_Module):
    """geotransformer/modules/geotransformer/superpoint_matching.py:7-83 (without the optional n2p-score gating,
    which model.py:308-311 does not use)."""

    def __init__(self, num_correspondences, dual_normalization=True, n2p_score_threshold=None):
        super().__init__()
        self.num_correspondences = num_correspondences
        self.dual_normalization = dual_normalization
        self.n2p_score_threshold = n2p_score_threshold

    def forward(self, ref_feats, src_feats, ref_masks=None, src_masks=None, ref_n2p_scores_c=None, src_n2p_scores_c=None):
        if ref_n2p_scores_c is not None:
            raise RuntimeError("n2p-score gating is not used by RDMNet.forward and is not implemented")
        if ref_masks is None:
            ref_masks = torch.ones(ref_feats.shape[0], dtype=torch.bool, device=ref_feats.device)
        if src_masks is None:
            src_masks = torch.ones(src_feats.shape[0], dtype=torch.bool, device=src_feats.device)
        return ops.coarse_matching(ref_feats, src_feats, ref_masks, src_masks, self.num_correspondences,
                                   self.dual_normalization)


class SuperPointTargetGenerator(_Module):
    """geotransformer/modules/geotransformer/superpoint_target.py:6-41: training-time sampler of ground-truth patch
    correspondences. The random draw stays numpy's global generator on the host (the reference's, :33: the same seed
    gives the same selection)."""

    def __init__(self, num_targets, overlap_threshold):
        super().__init__()
        self.num_targets = num_targets
        self.overlap_threshold = overlap_threshold

    @torch.no_grad()
    def forward(self, gt_corr_indices, gt_corr_overlaps):
        import numpy as np
        masks = torch.gt(gt_corr_overlaps, self.overlap_threshold)
        gt_corr_overlaps = gt_corr_overlaps[masks]
        gt_corr_indices = gt_corr_indices[masks]
        if gt_corr_indices.shape[0] > self.num_targets:
            sel = np.random.choice(np.arange(gt_corr_indices.shape[0]), self.num_targets, replace=False)
            sel = torch.from_numpy(sel).to(gt_corr_indices.device)
            gt_corr_indices = gt_corr_indices[sel]
            gt_corr_overlaps = gt_corr_overlaps[sel]
        return gt_corr_indices[:, 0], gt_corr_indices[:, 1], gt_corr_overlaps


class WeightedProcrustes(_Module):
    """geotransformer/modules/registration/procrustes.py:76-91."""

    def __init__(self, weight_thresh=0.0, eps=1e-5, return_transform=False):
        super().__init__()
        self.weight_thresh, self.eps, self.return_transform = weight_thresh, eps, return_transform

    def forward(self, src_points, tgt_points, weights=None):
        return ops.weighted_procrustes(src_points, tgt_points, weights, self.weight_thresh, self.eps, self.return_transform)


class LocalGlobalRegistration(_Module):
    """geotransformer/modules/geotransformer/local_global_registration.py:11-243 for the RDMNet configuration
    (k=1, mutual=False, use_dustbin=True, use_global_score=False, correspondence_limit=None)."""

    def __init__(self, k, acceptance_radius, mutual=True, confidence_threshold=0.05, use_dustbin=False,
                 use_global_score=False, correspondence_threshold=3, correspondence_limit=None, num_refinement_steps=5):
        super().__init__()
        if k != 1 or mutual or not use_dustbin or use_global_score or correspondence_limit is not None:
            raise ValueError("only k=1, mutual=False, use_dustbin=True, use_global_score=False, "
                             "correspondence_limit=None (experiments/config.py:152-161) is implemented")
        self.k, self.acceptance_radius, self.mutual = k, acceptance_radius, mutual
        self.confidence_threshold, self.use_dustbin, self.use_global_score = confidence_threshold, use_dustbin, use_global_score
        self.correspondence_threshold, self.correspondence_limit = correspondence_threshold, correspondence_limit
        self.num_refinement_steps = num_refinement_steps
        self.procrustes = WeightedProcrustes(return_transform=True)

    def forward(self, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, global_scores=None):
        """Reference signature: (B,K,3) x2, (B,K) x2 masks, (B,K+1,K+1) log scores. The batched knn tensors are
        treated as their own point tables (identity gather)."""
        b, k = ref_knn_masks.shape
        dev = score_mat.device
        ident = torch.arange(b * k, device=dev, dtype=torch.int64).view(b, k)
        rows = torch.arange(b, device=dev, dtype=torch.int64)
        ref_c, src_c, sc, T, _ = ops.local_global_registration(
            score_mat, ref_knn_points.reshape(-1, 3).contiguous(), src_knn_points.reshape(-1, 3).contiguous(), ident, ident,
            ref_knn_masks.to(torch.uint8).contiguous(), src_knn_masks.to(torch.uint8).contiguous(), rows, rows,
            self.acceptance_radius, self.correspondence_threshold, self.num_refinement_steps)
        return ref_c, src_c, sc, T
