"""2-D RoPE module with the reference's plugin interface.

Mirrors `cuRoPE2D` / `cuRoPE2D_func` (uniception/models/libs/croco/curope/curope2d.py:12-39): a
callable `(tokens[B,H,N,D], positions[B,N,2] int64) -> tokens`, applied IN PLACE (mark_dirty) by the
native kernel `uc_rope2d`, whose backward is the same kernel with `-F0`.  Attributes `base` and `F0`
are the ones the reference exposes (pos_embed.py:110-114) and are what lets the fused engine fold the
rotation into the QKV GEMM epilogue instead of calling this module.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class RoPE2DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, positions, base, F0=1.0):
        ctx.save_for_backward(positions)
        ctx.saved_base = base
        ctx.saved_F0 = F0
        ops.rope2d_(tokens, positions, base, F0)  # tokens: [B,N,H,D] view, rotated in place
        ctx.mark_dirty(tokens)
        return tokens

    @staticmethod
    def backward(ctx, grad_res):
        (positions,) = ctx.saved_tensors
        g = (grad_res if grad_res.stride(-1) == 1 else grad_res.contiguous()).clone()
        ops.rope2d_(g, positions, ctx.saved_base, -ctx.saved_F0)  # same kernel, -F0 (curope2d.py:24-28)
        return g, None, None, None


class RoPE2D(nn.Module):
    """Drop-in for `RoPE2D(freq=100.0, F0=1.0)` (libs/croco/pos_embed.py:109-155)."""

    def __init__(self, freq: float = 100.0, F0: float = 1.0):
        super().__init__()
        self.base = freq
        self.F0 = F0

    def forward(self, tokens, positions):
        assert tokens.size(3) % 2 == 0, "number of dimensions should be a multiple of two"
        assert positions.ndim == 3 and positions.shape[-1] == 2  # Batch, Seq, 2
        RoPE2DFunction.apply(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens


cuRoPE2D = RoPE2D


def fusable_rope(obj):
    """(base, F0) if `obj` is a RoPE2D-compatible positional encoding whose rotation can be fused
    into the GEMM epilogue, else None."""
    if obj is None:
        return None
    base, f0 = getattr(obj, "base", None), getattr(obj, "F0", None)
    if isinstance(base, (int, float)) and isinstance(f0, (int, float)):
        return float(base), float(f0)
    return None
