"""BASELINE.json configs[4]: "ViT-L encoder + DPT prediction_head dense depth, 518x518".

The reference has no single class for this pipeline; its users compose the modules by hand
(`CroCoIntermediateFeatureReturner` -> `DPTFeature` -> `DPTRegressionProcessor` -> `DepthAdaptor`; SURVEY.md 8c names this
composition as the oracle of the config: encoders/croco.py:260-327, prediction_heads/dpt.py:180-232, :285-311,
adaptors.py:233-257).  `ViTDPTDepth` is that composition as one module so that the whole model shares ONE flat parameter
pack: one bf16 cast per step, gradients of encoder and head in one buffer for the data-parallel all-reduce, the step
capturable as one CUDA graph.  The sub-modules keep the reference's constructor signatures and state-dict keys and can be
used on their own (tests/test_gpu_fullsize.py does).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn as nn

from .encoders import CroCoIntermediateFeatureReturner, feature_take_indices
from .params import ParamPack, get_pack
from .prediction_heads import AdaptorInput, DepthAdaptor, DPTFeature, DPTHead, DPTRegressionProcessor


class ViTDPTDepth(nn.Module):
    def __init__(self, name: str = "vit_dpt_depth", data_norm_type: str = "dust3r", img_size: Tuple[int, int] = (518, 518),
                 patch_size: int = 14, enc_embed_dim: int = 1024, enc_depth: int = 24, enc_num_heads: int = 16,
                 indices: Sequence[int] = (5, 11, 17, 23), feature_dim: int = 256, layer_dims: Sequence[int] = (96, 192, 384, 768),
                 depth_mode: str = "exp"):
        super().__init__()
        self.name = name
        self.encoder = CroCoIntermediateFeatureReturner(
            name=name + "_encoder", data_norm_type=data_norm_type, img_size=img_size, patch_size=patch_size,
            enc_embed_dim=enc_embed_dim, enc_depth=enc_depth, enc_num_heads=enc_num_heads, indices=list(indices),
            norm_intermediate=True, intermediates_only=True)
        self.dpt_feature_head = DPTFeature(patch_size=patch_size, hooks=[0, 1, 2, 3], input_feature_dims=[enc_embed_dim] * 4,
                                           layer_dims=list(layer_dims), feature_dim=feature_dim)
        self.dpt_regressor_head = DPTRegressionProcessor(input_feature_dim=feature_dim, output_dim=1)
        self.head = DPTHead(self.dpt_feature_head, self.dpt_regressor_head)  # alias, as factory/dust3r.py:178 does
        self.adaptor = DepthAdaptor(name="depth", mode=depth_mode)

    def pack(self) -> ParamPack:
        return get_pack(self)

    def forward(self, image: torch.Tensor, data_norm_type: str = "dust3r") -> torch.Tensor:
        """image [B,3,H,W] fp32 -> depth [B,1,H,W] fp32."""
        self.encoder._check_data_normalization_type(data_norm_type)
        B, _, H, W = image.shape
        p = self.encoder.patch_size
        h, w = H // p, W // p
        pk = self.pack()
        pk.refresh_bf16()
        take, _ = feature_take_indices(len(self.encoder.enc_blocks), self.encoder.indices)
        _, inter = self.encoder.forward_tokens(image, pk, "encoder.", take, self.encoder.norm_intermediate)
        y = self.head.forward_tokens(inter, B, h, w, (H, W))  # fp32 [B*H*W, 64], column 0 valid
        raw = y[:, :1].reshape(B, H, W, 1).permute(0, 3, 1, 2)
        return self.adaptor(AdaptorInput(adaptor_feature=raw, output_shape_hw=(H, W))).value
