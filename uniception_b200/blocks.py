"""Parameter containers with the reference's state-dict keys for the transformer blocks on the path.

`Mlp`, `Attention`, `Block` mirror libs/croco/blocks.py:64-161 (encoder flavour);
`CrossAttention`, `CrossAttentionBlock` mirror utils/transformer_blocks.py:260-386, :517-647.
`SelfAttentionBlock`, `LayerScale` mirror utils/transformer_blocks.py:389-514.
Built natively: the DUSt3R option set plus qk_norm, LayerScale (init_values) and the softmax-scaling flags; dropout,
stochastic depth and non-GELU activations raise NotImplementedError at construction instead of silently running something else.
Head dims other than 64 (multiples of 64) take the un-fused GEMM + row-softmax attention; the DiffAttention family lives in
diff_attention.py.

The fused whole-module engines (encoders.py / info_sharing.py) never call these modules' forward:
they read the parameters through a ParamPack.  The `forward` methods here exist so the blocks are
usable on their own with the reference's call signature; they run the same kernels through
`autograd_ops`.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Optional

import torch
import torch.nn as nn

from . import autograd_ops as A
from .engine import attn_scale


def _require(cond: bool, what: str):
    if not cond:
        raise NotImplementedError(f"uniception_b200: {what} is outside the B200 hot path (SURVEY.md 8f)")


class LayerScale(nn.Module):
    """utils/transformer_blocks.py:389-412 (parameter container; the fused engines apply gamma in `uc_layerscale_*`)."""

    def __init__(self, dim: int, init_values: float = 1e-5, inplace: bool = False):
        super().__init__()
        self.inplace = inplace
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return A.layer_scale(x, self.gamma)


def _head_norm(norm_layer, head_dim: int, qk_norm: bool) -> nn.Module:
    if not qk_norm:
        return nn.Identity()
    check_norm_layer(norm_layer)
    return norm_layer(head_dim)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, bias=True, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        _require(act_layer is nn.GELU, "a non-GELU activation")
        _require(drop == 0.0 and bias is True, "dropout / bias-free MLP")
        self.fc1 = nn.Linear(in_features, hidden_features, bias=True)
        self.act = nn.GELU()
        self.drop1 = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=True)
        self.drop2 = nn.Identity()

    def forward(self, x, residual=None):
        return A.mlp(x, self.fc1, self.fc2, residual)


class Attention(nn.Module):
    """Self-attention with both reference constructor flavours (`rope=` of libs/croco/blocks.py:90 and
    `custom_positional_encoding=` of utils/transformer_blocks.py:141-156)."""

    def __init__(self, dim, rope=None, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0, qk_norm=False,
                 custom_positional_encoding: Optional[Callable] = None, use_scalable_softmax=False, use_entropy_scaling=False,
                 base_token_count_for_entropy_scaling=444, entropy_scaling_growth_factor=1.4, norm_layer=nn.LayerNorm,
                 latent_attn_dim=None, **_ignored):
        super().__init__()
        if latent_attn_dim is not None:  # utils/transformer_blocks.py:178-199: q/k/v live in a latent width
            assert latent_attn_dim % num_heads == 0, "latent_attn_dim should be divisible by num_heads"
        else:
            assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.latent_attn = latent_attn_dim is not None
        width = latent_attn_dim if self.latent_attn else dim
        # head_dim 64: the fused tcgen05 attention kernels.  Other multiples of 64 (128 inside the DiffAttention family's
        # blocks): the un-fused GEMM + row-softmax path (autograd_ops.GeneralAttentionFn), without qk_norm
        _require((width // num_heads) % 64 == 0, f"head_dim {width // num_heads} (multiples of 64 only)")
        _require(width // num_heads == 64 or not qk_norm, "qk_norm with head_dim != 64")
        _require(attn_drop == 0.0 and proj_drop == 0.0, "attention dropout")
        # softmax scaling by the token count (utils/transformer_blocks.py:231-241) folds into the kernels' scale
        self.softmax_scaling = (use_scalable_softmax, use_entropy_scaling, base_token_count_for_entropy_scaling,
                                entropy_scaling_growth_factor) if (use_scalable_softmax or use_entropy_scaling) else None
        self.num_heads = num_heads
        self.head_dim = width // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, width * 3, bias=qkv_bias)
        self.q_norm = _head_norm(norm_layer, self.head_dim, qk_norm)  # registration order of transformer_blocks.py:194-202
        self.k_norm = _head_norm(norm_layer, self.head_dim, qk_norm)
        self.proj = nn.Linear(width, dim)
        self.rope = rope if rope is not None else custom_positional_encoding
        self.custom_positional_encoding = self.rope

    def forward(self, x, xpos=None, residual=None):
        B, N, _ = x.shape
        C = self.num_heads * self.head_dim  # the latent width with latent_attn_dim
        if self.rope is not None:
            assert xpos is not None, "Positions of tokens (xpos) are a required input when using custom positional encoding"
        qkv = A.linear(x, self.qkv.weight, self.qkv.bias)
        scale = attn_scale(self.softmax_scaling, N, self.head_dim)
        if self.head_dim != 64:
            qkv5 = qkv.reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
            q, k, v = qkv5[0], qkv5[1], qkv5[2]
            if self.rope is not None:  # the plugin may rotate in place (cuRoPE2D): give it private contiguous copies
                q = self.rope(q.clone(memory_format=torch.contiguous_format), xpos)
                k = self.rope(k.clone(memory_format=torch.contiguous_format), xpos)
            o = A.general_attention(q, k, v, scale).transpose(1, 2).reshape(B, N, C)
            return A.linear(o, self.proj.weight, self.proj.bias, residual=residual)
        if isinstance(self.q_norm, nn.LayerNorm):
            qn = A.head_norm(qkv, 0, C, self.q_norm)
            kv = torch.cat((A.head_norm(qkv, C, C, self.k_norm), qkv[..., 2 * C:]), dim=-1)
            o = A.attention(qn, kv, B, N, N, self.num_heads, q_off=0, k_off=0, v_off=C, qpos=xpos, kpos=xpos, rope=self.rope, scale=scale)
        else:
            o = A.attention(qkv, qkv, B, N, N, self.num_heads, q_off=0, k_off=C, v_off=2 * C, qpos=xpos, kpos=xpos, rope=self.rope,
                            scale=scale)
        return A.linear(o, self.proj.weight, self.proj.bias, residual=residual)


class Block(nn.Module):
    """libs/croco/blocks.py:133-161."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, rope=None):
        super().__init__()
        _require(drop_path == 0.0, "stochastic depth")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, rope=rope, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, xpos):
        x = self.attn(A.layer_norm(x, self.norm1), xpos, residual=x)  # residual add fused into the proj GEMM
        x = self.mlp(A.layer_norm(x, self.norm2), residual=x)
        return x


class CrossAttention(nn.Module):
    """utils/transformer_blocks.py:260-386."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, attn_drop=0.0, proj_drop=0.0,
                 norm_layer=nn.LayerNorm, custom_positional_encoding=None, use_scalable_softmax=False,
                 use_entropy_scaling=False, base_token_count_for_entropy_scaling=444, entropy_scaling_growth_factor=1.4, **_ignored):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        _require((dim // num_heads) % 64 == 0, f"head_dim {dim // num_heads} (multiples of 64 only)")
        _require(dim // num_heads == 64 or not qk_norm, "qk_norm with head_dim != 64")
        _require(attn_drop == 0.0 and proj_drop == 0.0, "attention dropout")
        self.softmax_scaling = (use_scalable_softmax, use_entropy_scaling, base_token_count_for_entropy_scaling,
                                entropy_scaling_growth_factor) if (use_scalable_softmax or use_entropy_scaling) else None
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.projq = nn.Linear(dim, dim, bias=qkv_bias)
        self.projk = nn.Linear(dim, dim, bias=qkv_bias)
        self.projv = nn.Linear(dim, dim, bias=qkv_bias)
        self.q_norm = _head_norm(norm_layer, self.head_dim, qk_norm)  # transformer_blocks.py:306-307
        self.k_norm = _head_norm(norm_layer, self.head_dim, qk_norm)
        self.proj = nn.Linear(dim, dim)
        self.custom_positional_encoding = custom_positional_encoding

    def forward(self, query, key, value, qpos=None, kpos=None, residual=None):
        B, Nq, C = query.shape
        Nk = key.shape[1]
        q = A.linear(query, self.projq.weight, self.projq.bias)
        k = A.linear(key, self.projk.weight, self.projk.bias)
        v = A.linear(value, self.projv.weight, self.projv.bias)
        if self.head_dim != 64:
            q4 = q.reshape(B, Nq, self.num_heads, self.head_dim).permute(0, 2, 1, 3)
            k4 = k.reshape(B, Nk, self.num_heads, self.head_dim).permute(0, 2, 1, 3)
            v4 = v.reshape(B, Nk, self.num_heads, self.head_dim).permute(0, 2, 1, 3)
            if self.custom_positional_encoding is not None:
                q4 = self.custom_positional_encoding(q4.clone(memory_format=torch.contiguous_format), qpos)
                k4 = self.custom_positional_encoding(k4.clone(memory_format=torch.contiguous_format), kpos)
            o = A.general_attention(q4, k4, v4, attn_scale(self.softmax_scaling, Nq, self.head_dim)).transpose(1, 2).reshape(B, Nq, C)
            return A.linear(o, self.proj.weight, self.proj.bias, residual=residual)
        if isinstance(self.q_norm, nn.LayerNorm):
            q, k = A.head_norm(q, 0, C, self.q_norm), A.head_norm(k, 0, C, self.k_norm)
        kv = torch.cat((k, v), dim=-1)
        o = A.attention(q, kv, B, Nq, Nk, self.num_heads, q_off=0, k_off=0, v_off=C, qpos=qpos, kpos=kpos,
                        rope=self.custom_positional_encoding, scale=attn_scale(self.softmax_scaling, Nq))
        return A.linear(o, self.proj.weight, self.proj.bias, residual=residual)


class CrossAttentionBlock(nn.Module):
    """utils/transformer_blocks.py:517-647 (registration order kept: norm1, attn, norm_y, norm2, cross_attn, norm3, mlp)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_norm=False, proj_drop=0.0, attn_drop=0.0,
                 init_values=None, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, mlp_layer=Mlp,
                 custom_positional_encoding=None, norm_cross_tokens=True, use_scalable_softmax=False,
                 use_entropy_scaling=False, base_token_count_for_entropy_scaling=444, entropy_scaling_growth_factor=1.4):
        super().__init__()
        _require(drop_path == 0.0, "stochastic depth")
        _require(mlp_layer is Mlp, "a custom mlp_layer")
        ls = (lambda: LayerScale(dim, init_values=init_values)) if init_values else nn.Identity
        common = dict(num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, attn_drop=attn_drop, proj_drop=proj_drop,
                      norm_layer=norm_layer,
                      custom_positional_encoding=custom_positional_encoding, use_scalable_softmax=use_scalable_softmax,
                      use_entropy_scaling=use_entropy_scaling,
                      base_token_count_for_entropy_scaling=base_token_count_for_entropy_scaling,
                      entropy_scaling_growth_factor=entropy_scaling_growth_factor)
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, **common)
        self.ls1 = ls()
        self.drop_path1 = nn.Identity()
        self.norm_y = norm_layer(dim) if norm_cross_tokens else nn.Identity()
        self.custom_positional_encoding = custom_positional_encoding
        self.norm2 = norm_layer(dim)
        self.cross_attn = CrossAttention(dim, **common)
        self.ls2 = ls()
        self.drop_path2 = nn.Identity()
        self.norm3 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=proj_drop)
        self.ls3 = ls()
        self.drop_path3 = nn.Identity()

    def forward(self, x, y, xpos=None, ypos=None):
        if self.custom_positional_encoding is not None:
            assert xpos is not None, "Positions of tokens (xpos) are a required input when using custom positional encoding"
            assert ypos is not None, "Positions of cross tokens (ypos) are a required input when using custom positional encoding"
        y_ = A.layer_norm(y, self.norm_y) if isinstance(self.norm_y, nn.LayerNorm) else y
        if isinstance(self.ls1, LayerScale):
            x = x + self.ls1(self.attn(A.layer_norm(x, self.norm1), xpos))
            x = x + self.ls2(self.cross_attn(A.layer_norm(x, self.norm2), y_, y_, xpos, ypos))
            return x + self.ls3(self.mlp(A.layer_norm(x, self.norm3)))
        x = self.attn(A.layer_norm(x, self.norm1), xpos, residual=x)
        x = self.cross_attn(A.layer_norm(x, self.norm2), y_, y_, xpos, ypos, residual=x)
        x = self.mlp(A.layer_norm(x, self.norm3), residual=x)
        return x


class SelfAttentionBlock(nn.Module):
    """utils/transformer_blocks.py:415-514 (registration order kept: norm1, attn, ls1, norm2, mlp, ls2)."""

    def __init__(self, dim, num_heads, latent_attn_dim=None, mlp_ratio=4.0, qkv_bias=False, qk_norm=False, proj_drop=0.0,
                 attn_drop=0.0, init_values=None, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm, mlp_layer=Mlp,
                 custom_positional_encoding=None, use_scalable_softmax=False, use_entropy_scaling=False,
                 base_token_count_for_entropy_scaling=444, entropy_scaling_growth_factor=1.4):
        super().__init__()
        _require(drop_path == 0.0, "stochastic depth")
        _require(mlp_layer is Mlp, "a custom mlp_layer")
        ls = (lambda: LayerScale(dim, init_values=init_values)) if init_values else nn.Identity
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, latent_attn_dim=latent_attn_dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, attn_drop=attn_drop, proj_drop=proj_drop,
                              norm_layer=norm_layer, custom_positional_encoding=custom_positional_encoding,
                              use_scalable_softmax=use_scalable_softmax, use_entropy_scaling=use_entropy_scaling,
                              base_token_count_for_entropy_scaling=base_token_count_for_entropy_scaling,
                              entropy_scaling_growth_factor=entropy_scaling_growth_factor)
        self.ls1 = ls()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = mlp_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=proj_drop)
        self.ls2 = ls()
        self.drop_path2 = nn.Identity()

    def forward(self, x, xpos=None):
        if isinstance(self.ls1, LayerScale):
            x = x + self.ls1(self.attn(A.layer_norm(x, self.norm1), xpos))
            return x + self.ls2(self.mlp(A.layer_norm(x, self.norm2)))
        x = self.attn(A.layer_norm(x, self.norm1), xpos, residual=x)
        return self.mlp(A.layer_norm(x, self.norm2), residual=x)


def check_norm_layer(norm_layer) -> None:
    """The engine implements nn.LayerNorm(dim, eps=1e-6) (encoders/croco.py:32); anything else is refused."""
    probe = norm_layer(8)
    _require(isinstance(probe, nn.LayerNorm) and abs(probe.eps - 1e-6) < 1e-12 and probe.elementwise_affine,
             f"norm_layer {norm_layer}")


DEFAULT_NORM = partial(nn.LayerNorm, eps=1e-6)
__all__ = ["Mlp", "Attention", "Block", "CrossAttention", "CrossAttentionBlock", "fusable_rope"]
