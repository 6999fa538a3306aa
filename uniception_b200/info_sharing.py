"""Drop-in multi-view cross-attention transformer (reference:
uniception/models/info_sharing/cross_attention_transformer.py, base.py).

Same constructor signatures, dataclass I/O, assertions and state-dict keys as
`MultiViewCrossAttentionTransformer` / `MultiViewCrossAttentionTransformerIFR`; per-view branches
hold separate weights (deep copies re-initialised, :148-160).  Forward = B200 engine (engine.decoder_fwd).
"""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass
from functools import partial
from typing import Callable, List, Optional, Union

import torch
import torch.nn as nn

from .checkpoints import load_checkpoint_file
from . import fused
from .blocks import CrossAttentionBlock, Mlp, SelfAttentionBlock, _require, check_norm_layer
from .encoders import IntermediateFeatureReturner, PositionGetter, feature_take_indices
from .params import ParamPack, get_pack
from .rope import RoPE2D, fusable_rope


# ---- dataclasses: info_sharing/base.py:14-97 ----
@dataclass
class InfoSharingInput:
    pass


@dataclass
class InfoSharingOutput:
    pass


@dataclass
class MultiViewTransformerInput(InfoSharingInput):
    features: List[torch.Tensor]  # per view [B, C_in, h, w]
    additional_input_tokens: Optional[torch.Tensor] = None
    additional_input_tokens_per_view: Optional[List[torch.Tensor]] = None


@dataclass
class MultiViewTransformerOutput(InfoSharingOutput):
    features: List[torch.Tensor]  # per view [B, dim, h, w]
    additional_token_features: Optional[torch.Tensor] = None
    additional_token_features_per_view: Optional[List[torch.Tensor]] = None


class UniCeptionInfoSharingBase(nn.Module):
    def __init__(self, name: str, size: Optional[str] = None, *args, **kwargs):
        super().__init__()
        self.name = name
        self.size = size


class MultiViewCrossAttentionTransformer(UniCeptionInfoSharingBase):
    "UniCeption Multi-View Cross-Attention Transformer on the B200 engine"

    def __init__(
        self,
        name: str,
        input_embed_dim: int,
        num_views: int,
        size: Optional[str] = None,
        depth: int = 12,
        dim: int = 768,
        num_heads: int = 12,
        mlp_ratio: float = 4.0,
        qkv_bias: bool = True,
        qk_norm: bool = False,
        proj_drop: float = 0.0,
        attn_drop: float = 0.0,
        init_values: Optional[float] = None,
        drop_path: float = 0.0,
        act_layer: nn.Module = nn.GELU,
        norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6),
        mlp_layer: nn.Module = Mlp,
        custom_positional_encoding: Callable = None,
        norm_cross_tokens: bool = True,
        use_scalable_softmax: bool = False,
        use_entropy_scaling: bool = False,
        base_token_count_for_entropy_scaling: int = 444,
        entropy_scaling_growth_factor: float = 1.4,
        pretrained_checkpoint_path: str = None,
        gradient_checkpointing: bool = False,
        *args,
        **kwargs,
    ):
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
        self.use_scalable_softmax = use_scalable_softmax
        self.use_entropy_scaling = use_entropy_scaling
        self.base_token_count_for_entropy_scaling = base_token_count_for_entropy_scaling
        self.entropy_scaling_growth_factor = entropy_scaling_growth_factor
        self.pretrained_checkpoint_path = pretrained_checkpoint_path
        self.gradient_checkpointing = gradient_checkpointing
        check_norm_layer(norm_layer)
        _require(qkv_bias, "qkv_bias=False")
        _require(dim // num_heads == 64, f"head_dim {dim // num_heads} in the fused decoder (only 64)")
        _require(not gradient_checkpointing,
                 "gradient_checkpointing (raises AttributeError in the reference too, cross_attention_transformer.py:163-165)")
        _require(custom_positional_encoding is None or fusable_rope(custom_positional_encoding) is not None,
                 "a positional encoding that is not RoPE2D-compatible (needs .base/.F0) in the fused decoder")

        if self.input_embed_dim != self.dim:
            self.proj_embed = nn.Linear(self.input_embed_dim, self.dim, bias=True)
        else:
            self.proj_embed = nn.Identity()

        blocks = nn.ModuleList(
            [CrossAttentionBlock(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm,
                                 proj_drop=proj_drop, attn_drop=attn_drop, init_values=init_values, drop_path=drop_path,
                                 act_layer=act_layer, norm_layer=norm_layer, mlp_layer=mlp_layer,
                                 custom_positional_encoding=custom_positional_encoding, norm_cross_tokens=norm_cross_tokens,
                                 use_scalable_softmax=use_scalable_softmax, use_entropy_scaling=use_entropy_scaling,
                                 base_token_count_for_entropy_scaling=base_token_count_for_entropy_scaling,
                                 entropy_scaling_growth_factor=entropy_scaling_growth_factor)
             for _ in range(depth)]
        )
        self.multi_view_branches = nn.ModuleList([blocks])
        for _ in range(1, self.num_views):
            self.multi_view_branches.append(deepcopy(blocks))
        self.norm = self.norm_layer(self.dim)
        if self.custom_positional_encoding is not None:
            self.position_getter = PositionGetter()
        self.initialize_weights()

        if self.pretrained_checkpoint_path is not None:
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

    # ---- engine entry ----
    def _cfg(self, B, h, w, take=(), norm_intermediate=True):
        fr = fusable_rope(self.custom_positional_encoding)
        sm = (self.use_scalable_softmax, self.use_entropy_scaling, self.base_token_count_for_entropy_scaling,
              self.entropy_scaling_growth_factor) if (self.use_scalable_softmax or self.use_entropy_scaling) else None
        return dict(B=B, h=h, w=w, depth=self.depth, heads=self.num_heads, rope_base=fr[0] if fr else None,
                    rope_f0=fr[1] if fr else 1.0, take=tuple(take), norm_intermediate=norm_intermediate,
                    has_proj_embed=isinstance(self.proj_embed, nn.Linear), has_norm_y=self.norm_cross_tokens,
                    softmax_scaling=sm)

    def forward_tokens(self, toks: List[torch.Tensor], B: int, h: int, w: int, pk: ParamPack, prefix: str, take=(),
                       norm_intermediate=True):
        """per-view tokens [B*h*w, C_in] -> (per-view normalised tokens bf16, [[per-view intermediate] per level])."""
        nv = len(toks)
        outs = fused.DecoderFn.apply(pk, prefix, self._cfg(B, h, w, take, norm_intermediate), nv, *toks, *pk.params.values())
        finals = list(outs[:nv])
        rest = outs[nv:]
        inter = [list(rest[l * nv:(l + 1) * nv]) for l in range(len(rest) // nv)]
        return finals, inter

    def _pack(self) -> ParamPack:
        pk = get_pack(self)
        pk.refresh_bf16()
        return pk

    def _check_input(self, model_input):
        assert len(model_input.features) == self.num_views, f"Expected {self.num_views} views, got {len(model_input.features)}"
        assert all(
            f.shape[1] == self.input_embed_dim for f in model_input.features
        ), f"All views must have input dimension {self.input_embed_dim}"
        assert all(f.ndim == 4 for f in model_input.features), "All views must have 4 dimensions (N, C, H, W)"
        if not model_input.features[0].is_cuda:
            raise RuntimeError("uniception_b200.MultiViewCrossAttentionTransformer runs on CUDA only (no CPU fallback)")

    def forward(self, model_input: MultiViewTransformerInput) -> MultiViewTransformerOutput:
        self._check_input(model_input)
        B, _, h, w = model_input.features[0].shape
        toks = [fused.NchwToNlcFn.apply(f) for f in model_input.features]
        finals, _ = self.forward_tokens(toks, B, h, w, self._pack(), "")
        return MultiViewTransformerOutput(features=[fused.NlcToNchwFn.apply(t, B, h, w) for t in finals])


class MultiViewCrossAttentionTransformerIFR(MultiViewCrossAttentionTransformer, IntermediateFeatureReturner):
    "Intermediate Feature Returner variant (cross_attention_transformer.py:278-505)"

    def __init__(self, name: str, input_embed_dim: int, num_views: int, size: Optional[str] = None, depth: int = 12,
                 dim: int = 768, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True, qk_norm: bool = False,
                 proj_drop: float = 0.0, attn_drop: float = 0.0, init_values: Optional[float] = None, drop_path: float = 0.0,
                 act_layer: nn.Module = nn.GELU, norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6),
                 mlp_layer: nn.Module = Mlp, custom_positional_encoding: Callable = None, norm_cross_tokens: bool = True,
                 use_scalable_softmax: bool = False, use_entropy_scaling: bool = False,
                 base_token_count_for_entropy_scaling: int = 444, entropy_scaling_growth_factor: float = 1.4,
                 pretrained_checkpoint_path: str = None, indices: Optional[Union[int, List[int]]] = None,
                 norm_intermediate: bool = True, intermediates_only: bool = False, gradient_checkpointing: bool = False,
                 *args, **kwargs):
        MultiViewCrossAttentionTransformer.__init__(
            self, name=name, input_embed_dim=input_embed_dim, num_views=num_views, size=size, depth=depth, dim=dim,
            num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm, proj_drop=proj_drop,
            attn_drop=attn_drop, init_values=init_values, drop_path=drop_path, act_layer=act_layer, norm_layer=norm_layer,
            mlp_layer=mlp_layer, custom_positional_encoding=custom_positional_encoding, norm_cross_tokens=norm_cross_tokens,
            use_scalable_softmax=use_scalable_softmax, use_entropy_scaling=use_entropy_scaling,
            base_token_count_for_entropy_scaling=base_token_count_for_entropy_scaling,
            entropy_scaling_growth_factor=entropy_scaling_growth_factor,
            pretrained_checkpoint_path=pretrained_checkpoint_path, gradient_checkpointing=gradient_checkpointing,
            *args, **kwargs)
        IntermediateFeatureReturner.__init__(self, indices=indices, norm_intermediate=norm_intermediate,
                                             intermediates_only=intermediates_only)

    def forward(self, model_input: MultiViewTransformerInput):
        self._check_input(model_input)
        B, _, h, w = model_input.features[0].shape
        take, _ = feature_take_indices(self.depth, self.indices)
        toks = [fused.NchwToNlcFn.apply(f) for f in model_input.features]
        finals, inter = self.forward_tokens(toks, B, h, w, self._pack(), "", take, self.norm_intermediate)
        inter_out = [MultiViewTransformerOutput(features=[fused.NlcToNchwFn.apply(t, B, h, w) for t in lvl]) for lvl in inter]
        if self.intermediates_only:
            return inter_out
        return MultiViewTransformerOutput(features=[fused.NlcToNchwFn.apply(t, B, h, w) for t in finals]), inter_out


def _sinusoid_table(n_position: int, d_hid: int, base: float = 10000.0) -> torch.Tensor:
    """global_attention_transformer.py:198-208 (`_get_sinusoid_encoding_table`): float64 numpy math, fp32 result."""
    import numpy as np

    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(base, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).float()


class MultiViewGlobalAttentionTransformer(UniCeptionInfoSharingBase):
    """UniCeption Multi-View Global-Attention Transformer on the B200 engine (global_attention_transformer.py:25-462):
    all views' tokens form one sequence of V*N tokens per batch element, `depth` SelfAttentionBlocks, final norm.
    Same constructor, state-dict keys (incl. the `view_pos_table` buffer) and I/O dataclasses as the reference.
    Additional input tokens (global and per-view) are supported.  Block flags built: qk_norm, LayerScale (init_values),
    scalable softmax / entropy scaling, gradient_checkpointing (per-block recompute).  Not built: dropout / stochastic depth."""

    ALTERNATING = False
    _PE_NON_REF_DEFAULT = True

    def __init__(self, name: str, input_embed_dim: int, distinguish_ref_and_non_ref_views: bool = True,
                 use_pe_for_non_reference_views: Optional[bool] = None, max_num_views_for_pe: int = 1000,
                 use_rand_idx_pe_for_non_reference_views: bool = True, size: Optional[str] = None, depth: int = 12,
                 dim: int = 768, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True, qk_norm: bool = False,
                 proj_drop: float = 0.0, attn_drop: float = 0.0, init_values: Optional[float] = None, drop_path: float = 0.0,
                 act_layer=nn.GELU, norm_layer=partial(nn.LayerNorm, eps=1e-6), mlp_layer=Mlp,
                 custom_positional_encoding=None, use_scalable_softmax: bool = False, use_entropy_scaling: bool = False,
                 base_token_count_for_entropy_scaling: int = 444, entropy_scaling_growth_factor: float = 1.4,
                 pretrained_checkpoint_path: Optional[str] = None, gradient_checkpointing: bool = False, *args, **kwargs):
        super().__init__(name=name, size=size, *args, **kwargs)
        if use_pe_for_non_reference_views is None:
            use_pe_for_non_reference_views = self._PE_NON_REF_DEFAULT
        check_norm_layer(norm_layer)
        if drop_path or proj_drop or attn_drop:
            raise NotImplementedError("uniception_b200: dropout / stochastic-depth block options (SURVEY.md 8f4)")
        _require(qkv_bias, "qkv_bias=False")  # the engine's Linear kernels always carry a bias (as the cross-attention class)
        _require(dim // num_heads == 64, f"head_dim {dim // num_heads} in the fused transformer (only 64)")
        self.qk_norm, self.init_values = qk_norm, init_values
        self.softmax_scaling = (use_scalable_softmax, use_entropy_scaling, base_token_count_for_entropy_scaling,
                                entropy_scaling_growth_factor) if (use_scalable_softmax or use_entropy_scaling) else None
        # per-block activation checkpointing (info_sharing/base.py:59-71): the engine keeps each block's input only and re-runs
        # the block's forward kernels in the backward pass
        self.gradient_checkpointing = gradient_checkpointing
        self.input_embed_dim = input_embed_dim
        self.distinguish_ref_and_non_ref_views = distinguish_ref_and_non_ref_views
        self.use_pe_for_non_reference_views = use_pe_for_non_reference_views
        self.max_num_views_for_pe = max_num_views_for_pe
        self.use_rand_idx_pe_for_non_reference_views = use_rand_idx_pe_for_non_reference_views
        self.depth, self.dim, self.num_heads, self.mlp_ratio, self.qkv_bias = depth, dim, num_heads, mlp_ratio, qkv_bias
        self.norm_layer = norm_layer
        self.pretrained_checkpoint_path = pretrained_checkpoint_path
        self.proj_embed = nn.Linear(input_embed_dim, dim, bias=True) if input_embed_dim != dim else nn.Identity()
        if isinstance(custom_positional_encoding, str):
            if custom_positional_encoding != "rope":
                raise ValueError(f"Unknown custom positional encoding: {custom_positional_encoding}")
            self.rope = RoPE2D(freq=100.0, F0=1.0)
            custom_positional_encoding = self.rope
        self.custom_positional_encoding = custom_positional_encoding
        if custom_positional_encoding is not None and fusable_rope(custom_positional_encoding) is None:
            raise NotImplementedError("uniception_b200: only RoPE2D positional encodings are fused")
        self.self_attention_blocks = nn.ModuleList(
            [SelfAttentionBlock(dim=dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_norm=qk_norm,
                                init_values=init_values, act_layer=act_layer, norm_layer=norm_layer, mlp_layer=mlp_layer,
                                custom_positional_encoding=custom_positional_encoding, use_scalable_softmax=use_scalable_softmax,
                                use_entropy_scaling=use_entropy_scaling,
                                base_token_count_for_entropy_scaling=base_token_count_for_entropy_scaling,
                                entropy_scaling_growth_factor=entropy_scaling_growth_factor)
             for _ in range(depth)])
        self.norm = norm_layer(dim)
        if distinguish_ref_and_non_ref_views:
            self.register_buffer("view_pos_table", _sinusoid_table(max_num_views_for_pe if use_pe_for_non_reference_views else 1, dim))
        self.apply(self._init_weights)
        if pretrained_checkpoint_path is not None:
            print(f"Loading pretrained multi-view attention transformer weights from {pretrained_checkpoint_path} ...")
            ckpt = load_checkpoint_file(pretrained_checkpoint_path)
            print(self.load_state_dict(ckpt["model"]))

    _init_weights = MultiViewCrossAttentionTransformer._init_weights

    def _pack(self) -> ParamPack:
        pk = get_pack(self)
        pk.refresh_bf16()
        return pk

    def _view_pe(self, nv: int) -> Optional[torch.Tensor]:
        """[V, dim] rows added to each view's tokens (global_attention_transformer.py:365-393)."""
        if not self.distinguish_ref_and_non_ref_views:
            return None
        pe = torch.zeros(nv, self.dim, device=self.view_pos_table.device)
        pe[0] = self.view_pos_table[0]
        if self.use_pe_for_non_reference_views and nv > 1:
            if self.use_rand_idx_pe_for_non_reference_views:
                idx = torch.randint(low=1, high=self.max_num_views_for_pe, size=(nv - 1,))
            else:
                idx = torch.arange(1, nv)
            pe[1:] = self.view_pos_table[idx.to(self.view_pos_table.device)]
        return pe

    def forward_sequence(self, x_in: torch.Tensor, B: int, nv: int, n_view: int, n_extra: int, h: int, w: int, pk: ParamPack,
                         prefix: str, take=(), norm_intermediate=True):
        """assembled token sequence [B*L, C_in] -> (final [B*L, dim], [tapped [B*L, dim]]); L = nv*n_view + n_extra."""
        fr = fusable_rope(self.custom_positional_encoding)
        cfg = dict(B=B, nv=nv, n_view=n_view, n_extra=n_extra, h=h, w=w, depth=self.depth, heads=self.num_heads,
                   rope_base=fr[0] if fr else None, rope_f0=fr[1] if fr else 1.0, alternating=self.ALTERNATING,
                   view_pe=self._view_pe(nv), has_proj_embed=isinstance(self.proj_embed, nn.Linear),
                   softmax_scaling=self.softmax_scaling, take=tuple(take), norm_intermediate=norm_intermediate,
                   recompute=bool(self.gradient_checkpointing) and torch.is_grad_enabled())
        outs = fused.MultiViewSelfAttnFn.apply(pk, prefix, cfg, x_in, *pk.params.values())
        return outs[0], list(outs[1:])

    def _check_input(self, model_input):
        feats = model_input.features
        assert len(feats) <= self.max_num_views_for_pe, f"Expected less than {self.max_num_views_for_pe} views, got {len(feats)}"
        assert all(f.shape[1] == self.input_embed_dim for f in feats), f"All views must have input dimension {self.input_embed_dim}"
        assert all(f.ndim == 4 for f in feats), "All views must have 4 dimensions (N, C, H, W)"
        B = feats[0].shape[0]
        per_view, extra = model_input.additional_input_tokens_per_view, model_input.additional_input_tokens
        if per_view is not None:  # global_attention_transformer.py:268-282
            assert len(per_view) == len(feats), \
                f"Number of additional token tensors ({len(per_view)}) must match number of views ({len(feats)})"
            assert all(t.ndim == 3 for t in per_view), "Additional tokens per view must have 3 dimensions (N, C, T)"
            assert all(t.shape[1] == self.input_embed_dim for t in per_view), \
                f"Additional tokens per view must have input dimension {self.input_embed_dim}"
            assert all(t.shape[0] == B for t in per_view), "Batch size mismatch for additional tokens per view"
        if extra is not None:  # :323-329
            assert extra.ndim == 3, "Additional tokens must have 3 dimensions (N, C, T)"
            assert extra.shape[1] == self.input_embed_dim, f"Additional tokens must have input dimension {self.input_embed_dim}"
            assert extra.shape[0] == B, "Batch size mismatch for additional tokens"
        if self.custom_positional_encoding is not None and (per_view is not None or extra is not None):  # :341-351
            raise ValueError("Custom positional encoding is not supported when additional_input_tokens or "
                             "additional_input_tokens_per_view are provided. Please set custom_positional_encoding=None "
                             "or remove additional tokens from the input.")
        if not feats[0].is_cuda:
            raise RuntimeError(f"uniception_b200.{type(self).__name__} runs on CUDA only (no CPU fallback)")

    def _assemble(self, model_input):
        """per-view maps (+ per-view / global additional tokens, [B, C, T]) -> bf16 [B*L, C_in] rows ordered (batch, [view,
        patch tokens then the view's extras], global extras) (global_attention_transformer.py:266-333)."""
        feats = model_input.features
        B, C, h, w = feats[0].shape
        N = h * w
        per_view, extra = model_input.additional_input_tokens_per_view, model_input.additional_input_tokens
        parts = []
        for v, f in enumerate(feats):
            parts.append(fused.NchwToNlcFn.apply(f).view(B, N, C))
            if per_view is not None:
                parts.append(per_view[v].permute(0, 2, 1).to(torch.bfloat16))
        n_view = N + (per_view[0].shape[2] if per_view is not None else 0)
        n_extra = 0
        if extra is not None:
            n_extra = extra.shape[2]
            parts.append(extra.permute(0, 2, 1).to(torch.bfloat16))
        x_in = torch.cat(parts, dim=1).reshape(B * (len(feats) * n_view + n_extra), C)
        return x_in, (B, len(feats), n_view, n_extra, h, w)

    def _split(self, y: torch.Tensor, layout, model_input) -> MultiViewTransformerOutput:
        """[B*L, dim] -> per-view maps (+ per-view / global additional token features [B, dim, T]) (:434-461)."""
        B, nv, n_view, n_extra, h, w = layout
        N = h * w
        y3 = y.view(B, nv * n_view + n_extra, -1)
        views = (y3[:, :nv * n_view] if n_extra else y3).reshape(B, nv, n_view, -1).unbind(1)
        feats = [fused.NlcToNchwFn.apply((t[:, :N] if n_view > N else t).reshape(B * N, -1), B, h, w) for t in views]
        per_view = None
        if model_input.additional_input_tokens_per_view is not None:
            per_view = [t[:, N:].permute(0, 2, 1).float().contiguous() for t in views]
        extra = y3[:, nv * n_view:].permute(0, 2, 1).float().contiguous() if model_input.additional_input_tokens is not None else None
        return MultiViewTransformerOutput(features=feats, additional_token_features=extra, additional_token_features_per_view=per_view)

    def forward(self, model_input: MultiViewTransformerInput) -> MultiViewTransformerOutput:
        self._check_input(model_input)
        x_in, layout = self._assemble(model_input)
        y, _ = self.forward_sequence(x_in, *layout, self._pack(), "")
        return self._split(y, layout, model_input)


class MultiViewGlobalAttentionTransformerIFR(MultiViewGlobalAttentionTransformer, IntermediateFeatureReturner):
    """Intermediate-feature-returner variant (global_attention_transformer.py:463-897): additionally returns the (optionally
    final-normed) token maps after the depths in `indices`."""

    def __init__(self, *args, indices: Optional[Union[int, List[int]]] = None, norm_intermediate: bool = True,
                 intermediates_only: bool = False, **kwargs):
        MultiViewGlobalAttentionTransformer.__init__(self, *args, **kwargs)
        IntermediateFeatureReturner.__init__(self, indices=indices, norm_intermediate=norm_intermediate,
                                             intermediates_only=intermediates_only)

    def forward(self, model_input: MultiViewTransformerInput):
        self._check_input(model_input)
        take, _ = feature_take_indices(self.depth, self.indices)
        x_in, layout = self._assemble(model_input)
        y, inter = self.forward_sequence(x_in, *layout, self._pack(), "", take, self.norm_intermediate)
        inter_out = [self._split(t, layout, model_input) for t in inter]
        if self.intermediates_only:
            return inter_out
        return self._split(y, layout, model_input), inter_out


class MultiViewAlternatingAttentionTransformer(MultiViewGlobalAttentionTransformer):
    """alternating_attention_transformer.py:22-500: even depths attend over all views' tokens, odd depths inside each view.
    (Only the reference view gets a view encoding by default: `use_pe_for_non_reference_views=False`, :30.)"""

    ALTERNATING = True
    _PE_NON_REF_DEFAULT = False


class MultiViewAlternatingAttentionTransformerIFR(MultiViewGlobalAttentionTransformerIFR):
    """alternating_attention_transformer.py:503-: intermediate-feature-returner variant of the alternating transformer."""

    ALTERNATING = True
    _PE_NON_REF_DEFAULT = False


# registry surface: info_sharing/__init__.py:23-37 (in-scope entries)
INFO_SHARING_CLASSES = {
    "cross_attention": (MultiViewCrossAttentionTransformer, MultiViewCrossAttentionTransformerIFR),
    "alternating_attention": (MultiViewAlternatingAttentionTransformer, MultiViewAlternatingAttentionTransformerIFR),
    "global_attention": (MultiViewGlobalAttentionTransformer, MultiViewGlobalAttentionTransformerIFR),
}
