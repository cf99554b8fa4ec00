"""`ffn_block`, `channel_normalization`, `Activation` with the reference's names, constructor arguments and
state_dict layout (reference modules/neural_net/common.py:185-220, 256-267).

The sub-modules are parameter containers: `ffn_block.forward` hands all of them to ONE fused CUDA tile
program (Linear -> per-row channel norm -> LeakyReLU); the intermediate tensors of the reference's
nn.Sequential never exist.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .constants import _EPS_, _LEAKY_RELU_NEG_SLOPE_


class channel_normalization(nn.Module):
    """y = std * (x - mean_row) / (std_row + eps) + mu with scalar learnable `std`, `mu`
    (reference common.py:208-220; std_row is the unbiased estimate)."""
    is_channel_normalization = True

    def __init__(self, eps: float = _EPS_):
        super().__init__()
        if eps != _EPS_:
            raise NotImplementedError('the CUDA kernels are built for eps = 1e-5')
        self.eps = eps
        self.mu = nn.Parameter(torch.zeros(1))
        self.std = nn.Parameter(torch.ones(1))

    def forward(self, x: torch.Tensor):
        raise NotImplementedError('channel_normalization only runs fused inside ffn_block (no stand-alone kernel)')


class layer_normalization(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError('layer_normalization mixes statistics across rows/frames; not implemented '
                                  '(the reference configuration uses channel_normalization)')


class group_normalization(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError('group_normalization mixes statistics across rows/frames; not implemented '
                                  '(the reference configuration uses channel_normalization)')


class Activation(nn.Module):
    """LeakyReLU(0.01) (reference common.py:256-267 with activation='leakyrelu')."""

    def __init__(self, activation: str = 'leakyrelu'):
        super().__init__()
        if activation != 'leakyrelu':
            raise NotImplementedError(f"activation '{activation}': only 'leakyrelu' (the reference yml) is implemented")
        self.kind = activation
        self.negative_slope = _LEAKY_RELU_NEG_SLOPE_

    def forward(self, x: torch.Tensor):
        raise NotImplementedError('Activation only runs fused inside ffn_block')


class ffn_block(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, activation: str,
                 norm_layer: Optional[str] = None, num_groups: Optional[int] = None):
        super().__init__()
        lin = nn.Linear(in_features=in_channels, out_features=out_channels, bias=True)
        act = Activation(activation)
        if norm_layer is not None:
            if norm_layer == 'channel_normalization':
                norm = channel_normalization()
            elif norm_layer == 'layer_normalization':
                norm = layer_normalization()
            elif norm_layer == 'group_normalization':
                norm = group_normalization(num_groups)
            else:
                raise ValueError(f'unknown norm_layer {norm_layer!r}')
            self.block = nn.Sequential(lin, norm, act)      # keys block.0.{weight,bias}, block.1.{mu,std}
        else:
            self.block = nn.Sequential(lin, act)

    def forward(self, x: torch.Tensor):
        from ._engine import apply_stack
        return apply_stack(x, [self])
