"""The DiffAttention family: differential attention layers, blocks and the differential multi-view
cross-attention transformer, with the reference's constructors, state-dict keys and I/O dataclasses.

Reference: `DiffAttention` / `DiffCrossAttention` / `DiffSelfAttentionBlock` / `DiffCrossAttentionBlock`
(utils/transformer_blocks.py:658-1031) and `DifferentialMultiViewCrossAttentionTransformer(IFR)`
(info_sharing/diff_cross_attention_transformer.py:22-588).

Shapes: a Diff layer built with `num_heads` heads splits q and k into 2 * num_heads heads of head_dim = dim / num_heads / 2 and
v into num_heads heads of 2 * head_dim; the transformer builds its blocks with num_heads // 2, so at the default dim 768 /
12 heads the differential attention pairs 64-wide q / k with 128-wide v and the plain self-attention inside every block has
128-wide heads.  Neither fits the fused head_dim-64 kernels: this family runs on the UN-FUSED attention
(`autograd_ops.GeneralAttentionFn`: uc_gemm scores -> uc_softmax_rows -> uc_gemm), Linear / LayerNorm / MLP on the same
kernels as everything else; the lambda combination and the RMS sub-layer norm are elementwise torch expressions (glue on a
[B, N, C] tensor).  Not a fused engine and not a benchmark path -- it exists so that the reference's registry name
"diff_cross_attention" resolves to working, parity-checked modules.

Kept quirks of the reference (parity, not taste): `DiffAttention.forward` reshapes the [B, H, N, 2d] result to [B, N, C]
WITHOUT transposing heads and tokens first (transformer_blocks.py:789), `DiffCrossAttention` does transpose (:921-922).
"""
from __future__ import annotations

import math
from copy import deepcopy
from functools import partial
from typing import Callable, List, Optional, Union

import torch
import torch.nn as nn

from . import autograd_ops as A
from .blocks import CrossAttentionBlock, Mlp, SelfAttentionBlock, _head_norm, _require
from .encoders import IntermediateFeatureReturner, PositionGetter, feature_take_indices
from .info_sharing import MultiViewTransformerInput, MultiViewTransformerOutput, UniCeptionInfoSharingBase


def lambda_init_fn(depth):
    """transformer_blocks.py:682-683."""
    return 0.8 - 0.6 * math.exp(-0.3 * depth)


class RMSNorm(nn.Module):
    """transformer_blocks.py:658-679."""

    def __init__(self, dim: int, eps: float = 1e-6, elementwise_affine=True, memory_efficient=False):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        if self.elementwise_affine:
            self.weight = nn.Parameter(torch.ones(dim))
        else:
            self.register_parameter("weight", None)

    def _norm(self, x):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps)

    def forward(self, x):
        output = self._norm(x.float()).type_as(x)
        if self.weight is not None:
            output = output * self.weight
        return output

    def extra_repr(self) -> str:
        return f"dim={self.dim}, eps={self.eps}, elementwise_affine={self.elementwise_affine}"


class _DiffBase(nn.Module):
    """lambda parameters + sub-layer norm shared by the two Diff layers (transformer_blocks.py:734-741, :869-876)."""

    def _init_diff(self, depth: int):
        self.lambda_init = lambda_init_fn(depth)
        for n in ("lambda_q1", "lambda_k1", "lambda_q2", "lambda_k2"):
            setattr(self, n, nn.Parameter(torch.zeros(self.head_dim, dtype=torch.float32).normal_(mean=0, std=0.1)))
        self.subln = RMSNorm(2 * self.head_dim, eps=1e-5, elementwise_affine=True)

    def _combine(self, attn1, attn2):
        """attn1 - lambda * attn2 -> RMS sub-norm -> x (1 - lambda_init), in fp32 (:779-787, :924-931)."""
        lambda_1 = torch.exp(torch.sum(self.lambda_q1 * self.lambda_k1, dim=-1).float())
        lambda_2 = torch.exp(torch.sum(self.lambda_q2 * self.lambda_k2, dim=-1).float())
        lambda_full = lambda_1 - lambda_2 + self.lambda_init
        attn = attn1.float() - lambda_full * attn2.float()
        return self.subln(attn) * (1 - self.lambda_init)

    def _norm_qk(self, q, k):
        if isinstance(self.q_norm, nn.LayerNorm):  # glue: LayerNorm over head_dim on [B, 2H, N, d]
            q, k = self.q_norm(q.float()), self.k_norm(k.float())
        return q, k


class DiffAttention(_DiffBase):
    "Differential Self-Attention Layer (transformer_blocks.py:686-795)"

    def __init__(self, dim: int, depth: int, num_heads: int = 8, qkv_bias: bool = False, qk_norm: bool = False, attn_drop: float = 0.0,
                 proj_drop: float = 0.0, norm_layer: nn.Module = nn.LayerNorm, custom_positional_encoding: Callable = None):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        _require(attn_drop == 0.0 and proj_drop == 0.0, "attention dropout")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads // 2
        _require(self.head_dim % 64 == 0, f"differential head_dim {self.head_dim} (multiples of 64 only)")
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = _head_norm(norm_layer, self.head_dim, qk_norm)
        self.k_norm = _head_norm(norm_layer, self.head_dim, qk_norm)
        self.attn_drop = nn.Identity()
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Identity()
        self.custom_positional_encoding = custom_positional_encoding
        self._init_diff(depth)

    def forward(self, x: torch.Tensor, xpos: torch.Tensor = None, residual=None) -> torch.Tensor:
        B, N, C = x.shape
        H, d = self.num_heads, self.head_dim
        qkv = A.linear(x, self.qkv.weight, self.qkv.bias).reshape(B, N, 3, H, 2 * d)
        q = qkv[:, :, 0].reshape(B, N, 2 * H, d).permute(0, 2, 1, 3)
        k = qkv[:, :, 1].reshape(B, N, 2 * H, d).permute(0, 2, 1, 3)
        v = qkv[:, :, 2].reshape(B, N, H, 2 * d).permute(0, 2, 1, 3)
        q, k = self._norm_qk(q, k)
        if self.custom_positional_encoding is not None:
            assert xpos is not None, "Positions of tokens (xpos) are a required input when using custom positional encoding"
            # the plugin may rotate in place (cuRoPE2D): give it private contiguous copies
            q = self.custom_positional_encoding(q.clone(memory_format=torch.contiguous_format), xpos)
            k = self.custom_positional_encoding(k.clone(memory_format=torch.contiguous_format), xpos)
        q1, q2 = q.chunk(2, dim=1)
        k1, k2 = k.chunk(2, dim=1)
        attn1 = A.general_attention(q1, k1, v, self.scale)
        attn2 = A.general_attention(q2, k2, v, self.scale)
        attn = self._combine(attn1, attn2)
        attn = attn.reshape(B, N, H * 2 * d)  # no head / token transpose here, as in the reference (:789)
        return A.linear(attn, self.proj.weight, self.proj.bias, residual=residual)


class DiffCrossAttention(_DiffBase):
    "Differential Cross-Attention Layer (transformer_blocks.py:798-938)"

    def __init__(self, dim: int, depth: int, num_heads: int = 8, qkv_bias: bool = False, qk_norm: bool = False, attn_drop: float = 0.0,
                 proj_drop: float = 0.0, norm_layer: nn.Module = nn.LayerNorm, custom_positional_encoding: Callable = None):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        _require(attn_drop == 0.0 and proj_drop == 0.0, "attention dropout")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads // 2
        _require(self.head_dim % 64 == 0, f"differential head_dim {self.head_dim} (multiples of 64 only)")
        self.scale = self.head_dim ** -0.5
        self.projq = nn.Linear(dim, dim, bias=qkv_bias)
        self.projk = nn.Linear(dim, dim, bias=qkv_bias)
        self.projv = nn.Linear(dim, dim, bias=qkv_bias)
        self.q_norm = _head_norm(norm_layer, self.head_dim, qk_norm)
        self.k_norm = _head_norm(norm_layer, self.head_dim, qk_norm)
        self.attn_drop = nn.Identity()
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Identity()
        self._init_diff(depth)
        self.custom_positional_encoding = custom_positional_encoding

    def lambda_init_fn(self, depth):
        return lambda_init_fn(depth)

    def forward(self, query, key, value, qpos=None, kpos=None, residual=None) -> torch.Tensor:
        B, Nq, C = query.shape
        Nk, Nv = key.shape[1], value.shape[1]
        H, d = self.num_heads, self.head_dim
        q = A.linear(query, self.projq.weight, self.projq.bias).reshape(B, Nq, 2 * H, d).permute(0, 2, 1, 3)
        k = A.linear(key, self.projk.weight, self.projk.bias).reshape(B, Nk, 2 * H, d).permute(0, 2, 1, 3)
        v = A.linear(value, self.projv.weight, self.projv.bias).reshape(B, Nv, H, 2 * d).permute(0, 2, 1, 3)
        q, k = self._norm_qk(q, k)
        if self.custom_positional_encoding is not None:
            assert qpos is not None, "Positions of queries (qpos) are a required input when using custom positional encoding"
            assert kpos is not None, "Positions of keys (kpos) are a required input when using custom positional encoding"
            q = self.custom_positional_encoding(q.clone(memory_format=torch.contiguous_format), qpos)
            k = self.custom_positional_encoding(k.clone(memory_format=torch.contiguous_format), kpos)
        q1, q2 = q.chunk(2, dim=1)
        k1, k2 = k.chunk(2, dim=1)
        attn1 = A.general_attention(q1, k1, v, self.scale).transpose(1, 2)  # B, Nq, H, 2d
        attn2 = A.general_attention(q2, k2, v, self.scale).transpose(1, 2)
        attn = self._combine(attn1, attn2).reshape(B, Nq, H * 2 * d)
        return A.linear(attn, self.proj.weight, self.proj.bias, residual=residual)


class DiffSelfAttentionBlock(SelfAttentionBlock):
    "Differential Self-Attention Block (transformer_blocks.py:941-986)"

    def __init__(self, dim: int, depth: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = False, qk_norm: bool = False,
                 proj_drop: float = 0.0, attn_drop: float = 0.0, init_values: Optional[float] = None, drop_path: float = 0.0,
                 act_layer: nn.Module = nn.GELU, norm_layer: nn.Module = nn.LayerNorm, mlp_layer: nn.Module = Mlp,
                 custom_positional_encoding: Callable = None):
        super().__init__(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=False, proj_drop=proj_drop,
                         attn_drop=attn_drop, init_values=init_values, drop_path=drop_path, act_layer=act_layer, norm_layer=norm_layer,
                         mlp_layer=mlp_layer, custom_positional_encoding=custom_positional_encoding)
        self.attn = DiffAttention(dim, depth, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, attn_drop=attn_drop,
                                  proj_drop=proj_drop, norm_layer=norm_layer, custom_positional_encoding=custom_positional_encoding)


class DiffCrossAttentionBlock(CrossAttentionBlock):
    """Differential Cross-Attention Block (transformer_blocks.py:989-1031): the plain self-attention of CrossAttentionBlock
    (head_dim = dim / num_heads, 128 at the defaults) followed by a DIFFERENTIAL cross-attention and the MLP."""

    def __init__(self, dim: int, depth: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = False, qk_norm: bool = False,
                 proj_drop: float = 0.0, attn_drop: float = 0.0, init_values: Optional[float] = None, drop_path: float = 0.0,
                 act_layer: nn.Module = nn.GELU, norm_layer: nn.Module = nn.LayerNorm, mlp_layer: nn.Module = Mlp,
                 custom_positional_encoding: Callable = None, norm_cross_tokens: bool = True):
        # the parent's own cross_attn is replaced below; build it without qk_norm so that head dims other than 64 construct
        super().__init__(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                         qk_norm=qk_norm and dim // num_heads == 64, proj_drop=proj_drop, attn_drop=attn_drop, init_values=init_values,
                         drop_path=drop_path, act_layer=act_layer, norm_layer=norm_layer, mlp_layer=mlp_layer,
                         custom_positional_encoding=custom_positional_encoding, norm_cross_tokens=norm_cross_tokens)
        _require(not qk_norm or dim // num_heads == 64, "qk_norm with a self-attention head_dim != 64")
        self.cross_attn = DiffCrossAttention(dim, depth, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, attn_drop=attn_drop,
                                             proj_drop=proj_drop, norm_layer=norm_layer,
                                             custom_positional_encoding=custom_positional_encoding)


class DifferentialMultiViewCrossAttentionTransformer(UniCeptionInfoSharingBase):
    """diff_cross_attention_transformer.py:22-259: per view a stack of DiffCrossAttentionBlocks (built with num_heads // 2);
    each view's block at depth k reads the OTHER views' tokens of depth k - 1; final LayerNorm; NCHW in / out."""

    def __init__(self, name: str, input_embed_dim: int, num_views: int, size: Optional[str] = None, depth: int = 12, dim: int = 768,
                 num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True, qk_norm: bool = False, proj_drop: float = 0.0,
                 attn_drop: float = 0.0, init_values: Optional[float] = None, drop_path: float = 0.0, act_layer: nn.Module = nn.GELU,
                 norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6), mlp_layer: nn.Module = Mlp,
                 custom_positional_encoding: Callable = None, norm_cross_tokens: bool = True, pretrained_checkpoint_path: str = None,
                 gradient_checkpointing: bool = False, *args, **kwargs):
        super().__init__(name=name, size=size, *args, **kwargs)
        self.input_embed_dim = input_embed_dim
        self.num_views = num_views
        self.depth = depth
        self.dim = dim
        self.num_heads = num_heads
        self.mlp_ratio = mlp_ratio
        self.qkv_bias = qkv_bias
        self.qk_norm = qk_norm
        self.proj_drop = proj_drop
        self.attn_drop = attn_drop
        self.init_values = init_values
        self.drop_path = drop_path
        self.act_layer = act_layer
        self.norm_layer = norm_layer
        self.mlp_layer = mlp_layer
        self.custom_positional_encoding = custom_positional_encoding
        self.norm_cross_tokens = norm_cross_tokens
        self.pretrained_checkpoint_path = pretrained_checkpoint_path
        self.gradient_checkpointing = gradient_checkpointing
        _require(not gradient_checkpointing,
                 "gradient_checkpointing (raises AttributeError in the reference too, diff_cross_attention_transformer.py:147-149)")
        if self.input_embed_dim != self.dim:
            self.proj_embed = nn.Linear(self.input_embed_dim, self.dim, bias=True)
        else:
            self.proj_embed = nn.Identity()
        assert num_heads % 2 == 0, "Number of heads must be divisible by 2 for differential cross-attention."
        blocks = nn.ModuleList(
            [DiffCrossAttentionBlock(depth=i, dim=dim, num_heads=num_heads // 2, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm,
                                     proj_drop=proj_drop, attn_drop=attn_drop, init_values=init_values, drop_path=drop_path,
                                     act_layer=act_layer, norm_layer=norm_layer, mlp_layer=mlp_layer,
                                     custom_positional_encoding=custom_positional_encoding, norm_cross_tokens=norm_cross_tokens)
             for i in range(depth)])
        self.multi_view_branches = nn.ModuleList([blocks])
        for _ in range(1, self.num_views):
            self.multi_view_branches.append(deepcopy(blocks))
        self.norm = self.norm_layer(self.dim)
        if self.custom_positional_encoding is not None:
            self.position_getter = PositionGetter()
        self.initialize_weights()
        if self.pretrained_checkpoint_path is not None:
            from .checkpoints import load_checkpoint_file

            print(f"Loading pretrained multi-view cross-attention transformer weights from {self.pretrained_checkpoint_path} ...")
            ckpt = load_checkpoint_file(self.pretrained_checkpoint_path)
            print(self.load_state_dict(ckpt["model"]))

    def initialize_weights(self):
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _check_input(self, model_input):
        assert len(model_input.features) == self.num_views, f"Expected {self.num_views} views, got {len(model_input.features)}"
        assert all(
            f.shape[1] == self.input_embed_dim for f in model_input.features
        ), f"All views must have input dimension {self.input_embed_dim}"
        assert all(f.ndim == 4 for f in model_input.features), "All views must have 4 dimensions (N, C, H, W)"
        if not model_input.features[0].is_cuda:
            raise RuntimeError("uniception_b200.DifferentialMultiViewCrossAttentionTransformer runs on CUDA only (no CPU fallback)")

    def _run(self, model_input, take=(), norm_intermediate=True):
        """Returns (per-view token tensors after the last block, [[per-view tokens] per taken depth], (B, h, w))."""
        self._check_input(model_input)
        B, _, h, w = model_input.features[0].shape
        toks = [f.permute(0, 2, 3, 1).reshape(B, h * w, self.input_embed_dim).contiguous() for f in model_input.features]
        if self.custom_positional_encoding is not None:
            pos = [self.position_getter(B, h, w, t.device) for t in toks]
        else:
            pos = [None] * self.num_views
        if isinstance(self.proj_embed, nn.Linear):
            toks = [A.linear(t, self.proj_embed.weight, self.proj_embed.bias) for t in toks]
        inter = []
        for k in range(self.depth):
            new = []
            for v, t in enumerate(toks):
                others = torch.cat([toks[i] for i in range(self.num_views) if i != v], dim=1)
                opos = torch.cat([pos[i] for i in range(self.num_views) if i != v], dim=1) if pos[v] is not None else None
                new.append(self.multi_view_branches[v][k](t, others, pos[v], opos))
            toks = new
            if k in take:
                inter.append([A.layer_norm(t, self.norm) if norm_intermediate else t for t in toks])
        return toks, inter, (B, h, w)

    def _to_nchw(self, t, B, h, w):
        return t.reshape(B, h, w, self.dim).permute(0, 3, 1, 2).float().contiguous()

    def forward(self, model_input: MultiViewTransformerInput) -> MultiViewTransformerOutput:
        toks, _, (B, h, w) = self._run(model_input)
        return MultiViewTransformerOutput(features=[self._to_nchw(A.layer_norm(t, self.norm), B, h, w) for t in toks])


class DifferentialMultiViewCrossAttentionTransformerIFR(DifferentialMultiViewCrossAttentionTransformer, IntermediateFeatureReturner):
    "Intermediate Feature Returner variant (diff_cross_attention_transformer.py:262-507)"

    def __init__(self, name: str, input_embed_dim: int, num_views: int, size: Optional[str] = None, depth: int = 12, dim: int = 768,
                 num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True, qk_norm: bool = False, proj_drop: float = 0.0,
                 attn_drop: float = 0.0, init_values: Optional[float] = None, drop_path: float = 0.0, act_layer: nn.Module = nn.GELU,
                 norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6), mlp_layer: nn.Module = Mlp,
                 custom_positional_encoding: Callable = None, norm_cross_tokens: bool = True, pretrained_checkpoint_path: str = None,
                 indices: Optional[Union[int, List[int]]] = None, norm_intermediate: bool = True, intermediates_only: bool = False,
                 gradient_checkpointing: bool = False, *args, **kwargs):
        DifferentialMultiViewCrossAttentionTransformer.__init__(
            self, name=name, input_embed_dim=input_embed_dim, num_views=num_views, size=size, depth=depth, dim=dim, num_heads=num_heads,
            mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm, proj_drop=proj_drop, attn_drop=attn_drop, init_values=init_values,
            drop_path=drop_path, act_layer=act_layer, norm_layer=norm_layer, mlp_layer=mlp_layer,
            custom_positional_encoding=custom_positional_encoding, norm_cross_tokens=norm_cross_tokens,
            pretrained_checkpoint_path=pretrained_checkpoint_path, gradient_checkpointing=gradient_checkpointing, *args, **kwargs)
        IntermediateFeatureReturner.__init__(self, indices=indices, norm_intermediate=norm_intermediate,
                                             intermediates_only=intermediates_only)

    def forward(self, model_input: MultiViewTransformerInput):
        take, _ = feature_take_indices(self.depth, self.indices)
        toks, inter, (B, h, w) = self._run(model_input, take, self.norm_intermediate)
        inter_out = [MultiViewTransformerOutput(features=[self._to_nchw(t, B, h, w) for t in lvl]) for lvl in inter]
        if self.intermediates_only:
            return inter_out
        return MultiViewTransformerOutput(features=[self._to_nchw(A.layer_norm(t, self.norm), B, h, w) for t in toks]), inter_out


# registered under the reference's name (info_sharing/__init__.py:25-28)
from . import info_sharing as _info  # noqa: E402

_info.INFO_SHARING_CLASSES["diff_cross_attention"] = (DifferentialMultiViewCrossAttentionTransformer,
                                                      DifferentialMultiViewCrossAttentionTransformerIFR)
