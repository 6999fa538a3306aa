"""Drop-in CroCo / DUSt3R ViT encoder (reference: uniception/models/encoders/croco.py, base.py).

Same constructor signature, attributes, dataclass I/O, assertions and state-dict keys as
`CroCoEncoder` / `CroCoIntermediateFeatureReturner`; the forward pass is the hand-scheduled B200
engine (patch gather -> tcgen05 GEMMs with fused bias / RoPE / GELU / residual epilogues -> fused
attention -> LayerNorm kernels).  CUDA only: there is no CPU or library fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import partial
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from .checkpoints import load_checkpoint_file
from . import fused
from .blocks import Block, _require, check_norm_layer
from .params import ParamPack, get_pack
from .rope import RoPE2D, fusable_rope


# ---- dataclasses: encoders/base.py:15-118 ----
@dataclass
class EncoderInput:
    data_norm_type: str


@dataclass
class EncoderOutput:
    pass


@dataclass
class ViTEncoderInput(EncoderInput):
    image: torch.Tensor  # [B, C, H, W]
    # optional attribute `true_shape` [B, 2] (height, width), attached by the caller as in the reference
    # (croco.py:160-165 reads it with hasattr): used by the ManyAR patch-embed


@dataclass
class ViTEncoderOutput(EncoderOutput):
    features: torch.Tensor  # [B, enc_embed_dim, H/p, W/p]
    registers: Optional[torch.Tensor] = None


def feature_take_indices(num_features: int, indices: Optional[Union[int, List[int]]] = None) -> Tuple[List[int], int]:
    """utils/intermediate_feature_return.py:47-85."""
    if indices is None:
        indices = num_features
    if isinstance(indices, int):
        assert 0 < indices <= num_features, f"last-n ({indices}) is out of range (1 to {num_features})"
        take_indices = [num_features - indices + i for i in range(indices)]
    else:
        take_indices = []
        for i in indices:
            idx = num_features + i if i < 0 else i
            assert 0 <= idx < num_features, f"feature index {idx} is out of range (0 to {num_features - 1})"
            take_indices.append(idx)
    return take_indices, max(take_indices)


class IntermediateFeatureReturner:
    """utils/intermediate_feature_return.py:19-43."""

    def __init__(self, indices=None, norm_intermediate: bool = True, stop_early: bool = False, intermediates_only: bool = True):
        self.indices = indices
        self.norm_intermediate = norm_intermediate
        self.stop_early = stop_early
        self.intermediates_only = intermediates_only


class PositionGetter:
    """(y, x) patch positions, int64 [B, h*w, 2] (libs/croco/patch_embed.py:19-31); cached per (h, w, device)."""

    def __init__(self):
        self.cache_positions = {}

    def __call__(self, b, h, w, device):
        key = (h, w, str(device))
        if key not in self.cache_positions:
            x = torch.arange(w, device=device)
            y = torch.arange(h, device=device)
            self.cache_positions[key] = torch.cartesian_prod(y, x)
        return self.cache_positions[key].view(1, h * w, 2).expand(b, -1, 2).clone()


class PatchEmbedDust3R(nn.Module):
    """Parameter container for the patch-embedding conv (libs/croco/patch_embed.py:34-82)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = True
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.Identity()
        self.position_getter = PositionGetter()

    def _init_weights(self):
        w = self.proj.weight.data
        torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))


class UniCeptionViTEncoderBase(nn.Module):
    """encoders/base.py:43-151 (name/size/data_norm_type/patch_size attributes + norm-type check)."""

    def __init__(self, name: str, data_norm_type: str, patch_size: int, size: Optional[str] = None,
                 gradient_checkpointing: bool = False, *args, **kwargs):
        super().__init__()
        self.name = name
        self.size = size
        self.data_norm_type = data_norm_type
        self.patch_size = patch_size
        self.gradient_checkpointing = gradient_checkpointing  # accepted, unused (as in the reference, croco.py:119-127)

    def _check_data_normalization_type(self, data_norm_type: str):
        assert (
            data_norm_type == self.data_norm_type
        ), f"Input normalization type {data_norm_type} does not match the encoder's normalization type {self.data_norm_type}."


class ManyAR_PatchEmbed(PatchEmbedDust3R):
    """Parameter container of libs/croco/patch_embed.py:85-127 (same `proj.*` keys): all images of a batch are stored in
    landscape orientation; `true_shape` says which samples are really portrait (their patches / positions are taken from
    the transposed image).  The arithmetic lives in engine.encoder_fwd(portrait=...)."""

    @staticmethod
    def portrait_flags(image: torch.Tensor, true_shape: Optional[torch.Tensor]):
        B, _, H, W = image.shape
        assert W >= H, f"img should be in landscape mode, but got W={W} H={H}"
        if true_shape is None:
            return None
        assert tuple(true_shape.shape) == (B, 2), f"true_shape has the wrong shape={tuple(true_shape.shape)}"
        height, width = true_shape.T
        return (width < height).tolist()


class CroCoEncoder(UniCeptionViTEncoderBase):
    "UniCeption CroCov2 Encoder on the B200 engine"

    def __init__(
        self,
        name: str,
        data_norm_type: str,
        patch_embed_cls: str = "PatchEmbedDust3R",
        img_size: Union[int, Tuple[int, int]] = (224, 224),
        patch_size: int = 16,
        enc_embed_dim: int = 1024,
        enc_depth: int = 24,
        enc_num_heads: int = 16,
        mlp_ratio: int = 4,
        norm_layer: Callable = partial(nn.LayerNorm, eps=1e-6),
        pos_embed: str = "RoPE100",
        pretrained_checkpoint_path: str = None,
        override_checkpoint_attributes: bool = False,
        *args,
        **kwargs,
    ):
        super().__init__(name=name, data_norm_type=data_norm_type, patch_size=patch_size, *args, **kwargs)
        self.patch_embed_cls = patch_embed_cls
        self.img_size = img_size
        self.enc_embed_dim = enc_embed_dim
        self.enc_depth = enc_depth
        self.enc_num_heads = enc_num_heads
        self.mlp_ratio = mlp_ratio
        self.norm_layer = norm_layer
        self.pretrained_checkpoint_path = pretrained_checkpoint_path
        self.override_checkpoint_attributes = override_checkpoint_attributes
        check_norm_layer(norm_layer)
        if patch_embed_cls not in ("PatchEmbedDust3R", "PatchEmbedCroCo", "ManyAR_PatchEmbed"):
            raise NotImplementedError(f"uniception_b200: unknown patch_embed_cls {patch_embed_cls!r}")

        self.pos_embed = pos_embed
        if pos_embed.startswith("RoPE"):  # eg RoPE100 (croco.py:79-85)
            self.enc_pos_embed = None
            self.dec_pos_embed = None
            freq = float(pos_embed[len("RoPE"):])
            self.rope = RoPE2D(freq=freq)
        else:
            raise NotImplementedError("Unknown pos_embed " + pos_embed)

        _require(enc_embed_dim // enc_num_heads == 64, f"head_dim {enc_embed_dim // enc_num_heads} in the fused encoder (only 64)")
        pe_cls = ManyAR_PatchEmbed if patch_embed_cls == "ManyAR_PatchEmbed" else PatchEmbedDust3R
        self.patch_embed = pe_cls(img_size, patch_size, 3, enc_embed_dim)
        self.enc_blocks = nn.ModuleList(
            [Block(enc_embed_dim, enc_num_heads, mlp_ratio, qkv_bias=True, norm_layer=norm_layer, rope=self.rope)
             for _ in range(enc_depth)]
        )
        self.enc_norm = norm_layer(enc_embed_dim)
        self.initialize_weights()

        if pretrained_checkpoint_path:
            print(f"Loading pretrained CroCo checkpoint from {pretrained_checkpoint_path}")
            ckpt = load_checkpoint_file(pretrained_checkpoint_path)
            print(self.load_state_dict(ckpt["model"]))
            if not override_checkpoint_attributes:
                assert (
                    data_norm_type == ckpt["data_norm_type"]
                ), f"Data normalization type {data_norm_type} does not match the checkpoint {ckpt['data_norm_type']}."
                assert (
                    patch_embed_cls == ckpt["patch_embed_cls"]
                ), f"Patch embedding class {patch_embed_cls} does not match the checkpoint {ckpt['patch_embed_cls']}."

    # ---- init: croco.py:129-145 ----
    def initialize_weights(self):
        self.patch_embed._init_weights()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- engine entry: tokens in / tokens out (used by the fused DUSt3R model) ----
    def _cfg(self, take=(), norm_intermediate=True, portrait=None):
        fr = fusable_rope(self.rope)
        return dict(depth=self.enc_depth, heads=self.enc_num_heads, patch=self.patch_size,
                    rope_base=fr[0] if fr else None, rope_f0=fr[1] if fr else 1.0, take=tuple(take),
                    norm_intermediate=norm_intermediate, portrait=portrait)

    def _portrait(self, encoder_input) -> Optional[list]:
        """ManyAR patch-embed: per-sample portrait flags from `encoder_input.true_shape` (croco.py:160-168)."""
        if not isinstance(self.patch_embed, ManyAR_PatchEmbed):
            return None
        return ManyAR_PatchEmbed.portrait_flags(encoder_input.image, getattr(encoder_input, "true_shape", None))

    def forward_tokens(self, image: torch.Tensor, pk: ParamPack, prefix: str, take=(), norm_intermediate=True, portrait=None):
        """image [B,3,H,W] -> (normalised tokens bf16 [B*N, C], [intermediate tokens...])."""
        B, Cin, H, W = image.shape
        assert H % self.patch_size == 0, f"Input image height ({H}) is not a multiple of patch size ({self.patch_size})."
        assert W % self.patch_size == 0, f"Input image width ({W}) is not a multiple of patch size ({self.patch_size})."
        if not image.is_cuda:
            raise RuntimeError("uniception_b200.CroCoEncoder runs on CUDA only (no CPU fallback)")
        outs = fused.EncoderFn.apply(image, pk, prefix, self._cfg(take, norm_intermediate, portrait), *pk.params.values())
        return outs[0], list(outs[1:])

    def _pack(self) -> ParamPack:
        pk = get_pack(self)
        pk.refresh_bf16()
        return pk

    def forward(self, encoder_input: ViTEncoderInput) -> ViTEncoderOutput:
        self._check_data_normalization_type(encoder_input.data_norm_type)
        B, _, H, W = encoder_input.image.shape
        tok, _ = self.forward_tokens(encoder_input.image, self._pack(), "", portrait=self._portrait(encoder_input))
        feats = fused.NlcToNchwFn.apply(tok, B, H // self.patch_size, W // self.patch_size)
        return ViTEncoderOutput(features=feats)


class CroCoIntermediateFeatureReturner(CroCoEncoder, IntermediateFeatureReturner):
    "Intermediate Feature Returner for the CroCo encoder (croco.py:185-327)"

    def __init__(self, name: str, data_norm_type: str, patch_embed_cls: str = "PatchEmbedDust3R",
                 img_size=(224, 224), patch_size: int = 16, enc_embed_dim: int = 1024, enc_depth: int = 24,
                 enc_num_heads: int = 16, mlp_ratio: int = 4, norm_layer: Callable = partial(nn.LayerNorm, eps=1e-6),
                 pos_embed: str = "RoPE100", pretrained_checkpoint_path: str = None,
                 indices: Optional[Union[int, List[int]]] = None, norm_intermediate: bool = True, stop_early: bool = False,
                 intermediates_only: bool = True, *args, **kwargs):
        CroCoEncoder.__init__(self, name=name, data_norm_type=data_norm_type, patch_embed_cls=patch_embed_cls,
                              img_size=img_size, patch_size=patch_size, enc_embed_dim=enc_embed_dim, enc_depth=enc_depth,
                              enc_num_heads=enc_num_heads, mlp_ratio=mlp_ratio, norm_layer=norm_layer, pos_embed=pos_embed,
                              pretrained_checkpoint_path=pretrained_checkpoint_path, *args, **kwargs)
        IntermediateFeatureReturner.__init__(self, indices=indices, norm_intermediate=norm_intermediate,
                                             stop_early=stop_early, intermediates_only=intermediates_only)

    def forward(self, encoder_input: ViTEncoderInput):
        self._check_data_normalization_type(encoder_input.data_norm_type)
        B, _, H, W = encoder_input.image.shape
        h, w = H // self.patch_size, W // self.patch_size
        take, _ = feature_take_indices(len(self.enc_blocks), self.indices)
        tok, inter = self.forward_tokens(encoder_input.image, self._pack(), "", take, self.norm_intermediate,
                                         portrait=self._portrait(encoder_input))
        inter = [ViTEncoderOutput(features=fused.NlcToNchwFn.apply(t, B, h, w)) for t in inter]
        if self.intermediates_only:
            return inter
        return ViTEncoderOutput(features=fused.NlcToNchwFn.apply(tok, B, h, w)), inter


# ---- registry surface: encoders/__init__.py:37-140 (only the in-scope entries) ----
ENCODER_CONFIGS = {
    "croco": {"class": CroCoEncoder, "intermediate_feature_returner_class": CroCoIntermediateFeatureReturner,
              "supported_models": ["CroCov2", "DUSt3R", "MASt3R"]},
}


def encoder_factory(encoder_str: str, **kwargs) -> nn.Module:
    if encoder_str not in ENCODER_CONFIGS:
        raise ValueError(f"Unknown encoder: {encoder_str}. For valid encoder_str options, see list(ENCODER_CONFIGS.keys())")
    return ENCODER_CONFIGS[encoder_str]["class"](**kwargs)


def feature_returner_encoder_factory(encoder_str: str, **kwargs) -> nn.Module:
    if encoder_str not in ENCODER_CONFIGS:
        raise ValueError(f"Unknown encoder: {encoder_str}. For valid encoder_str options, see list(ENCODER_CONFIGS.keys())")
    return ENCODER_CONFIGS[encoder_str]["intermediate_feature_returner_class"](**kwargs)
