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


class KPConv(nn.Module):
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


class GroupNorm(nn.Module):
    """geotransformer/modules/kpconv/modules.py:33-50 (statistics over the whole stacked (N, C/G) slab)."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups, self.num_channels = num_groups, num_channels
        self.norm = nn.GroupNorm(num_groups, num_channels)  # parameter holder: keys norm.weight / norm.bias

    def forward(self, x, residual=None, act=0):
        return ops.group_norm(x, self.norm.weight, self.norm.bias, self.num_groups, residual, act, 0.1, self.norm.eps)


class UnaryBlock(nn.Module):
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


class LastUnaryBlock(nn.Module):
    """kpconv/modules.py:86-101."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x):
        return ops.linear(x, self.mlp.weight, self.mlp.bias)


class ConvBlock(nn.Module):
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


class ResidualBlock(nn.Module):
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
