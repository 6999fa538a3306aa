"""Drop-in linear prediction head and output adaptors (reference:
uniception/models/prediction_heads/{linear,adaptors,base}.py).

`LinearFeature` keeps the reference's `linear.{weight[out*p*p, C, 1, 1], bias}` parameters; its 1x1
conv is the tcgen05 GEMM.  In the fused DUSt3R model the pixel-shuffle, the pointmap/confidence
adaptor and the BCHW->BHWC permutes are ONE kernel (`uc_head_post_fwd/bwd`, see
`LinearFeature.forward_fused`).  The stand-alone adaptor modules below keep the reference call
signature for composing other models; they are thin elementwise expressions on a [B,C,H,W] map.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch
import torch.nn as nn

from . import fused
from .autograd_ops import LinearFn
from .params import ParamPack, get_pack


# ---- dataclasses: prediction_heads/base.py:14-104 ----
@dataclass
class PredictionHeadInput:
    last_feature: torch.Tensor  # [B, C, h, w]


@dataclass
class PredictionHeadLayeredInput:
    list_features: List[torch.Tensor]
    target_output_shape: Tuple[int, int]


@dataclass
class PixelTaskOutput:
    decoded_channels: torch.Tensor  # [B, C_out, H, W]


@dataclass
class AdaptorInput:
    adaptor_feature: torch.Tensor
    output_shape_hw: Tuple[int, int]


@dataclass
class AdaptorOutput:
    value: torch.Tensor


@dataclass
class RegressionAdaptorOutput:
    value: torch.Tensor


@dataclass
class RegressionWithConfidenceAdaptorOutput:
    value: torch.Tensor
    confidence: torch.Tensor


class LinearFeature(nn.Module):
    """prediction_heads/linear.py:15-84."""

    def __init__(self, input_feature_dim: int, output_dim: int, patch_size: int, pretrained_checkpoint_path: str = None,
                 *args, **kwargs):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.output_dim = output_dim
        self.patch_size = patch_size
        self.pretrained_checkpoint_path = pretrained_checkpoint_path
        self.linear = nn.Conv2d(in_channels=input_feature_dim, out_channels=output_dim * (patch_size ** 2), kernel_size=1,
                                stride=1, padding=0, bias=True)
        if self.pretrained_checkpoint_path is not None:
            print(f"Loading pretrained linear dense feature head from {self.pretrained_checkpoint_path}")
            ckpt = torch.load(self.pretrained_checkpoint_path, weights_only=False)
            print(self.load_state_dict(ckpt["model"]))

    def forward(self, feature_input: PredictionHeadInput) -> PixelTaskOutput:
        x = feature_input.last_feature
        assert x.shape[1] == self.input_feature_dim, f"Input feature dimension mismatch: {x.shape[1]} != {self.input_feature_dim}"
        if not x.is_cuda:
            raise RuntimeError("uniception_b200.LinearFeature runs on CUDA only (no CPU fallback)")
        B, C, h, w = x.shape
        p = self.patch_size
        tok = fused.NchwToNlcFn.apply(x)                                    # [B*h*w, C] bf16
        y = LinearFn.apply(tok, self.linear.weight, self.linear.bias, None)  # [B*h*w, out*p*p]
        # pixel_shuffle as a pure index op (linear.py:81-82): out[b,c,p*h+i,p*w+j] = y[b,h,w,c*p*p+i*p+j]
        y = y.float().view(B, h, w, self.output_dim, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, self.output_dim, h * p, w * p)
        return PixelTaskOutput(decoded_channels=y)

    def forward_fused(self, tok: torch.Tensor, B: int, h: int, w: int, pk: ParamPack, prefix: str, conf_min: float, conf_max: float):
        """tokens -> (pts3d [B,H,W,3], conf [B,H,W,1]) through the fused head kernel (output_dim must be 4)."""
        assert self.output_dim == 4
        cfg = dict(B=B, h=h, w=w, patch=self.patch_size, conf_min=conf_min, conf_max=conf_max)
        return fused.LinearHeadFn.apply(tok, pk, prefix, cfg, *pk.params.values())


class UniCeptionAdaptorBase(nn.Module):
    def __init__(self, name: str, required_channels: int, *args, **kwargs):
        super().__init__()
        self.name = name
        self.required_channels = required_channels


class PointMapAdaptor(UniCeptionAdaptorBase):
    """prediction_heads/adaptors.py:299-355."""

    def __init__(self, name: str, mode: str, vmin: float = -math.inf, vmax: float = math.inf, *args, **kwargs):
        super().__init__(name, required_channels=3)
        self.mode = mode
        self.vmin = vmin
        self.vmax = vmax
        self.no_bounds = (vmin == -float("inf")) and (vmax == float("inf"))

    def forward(self, adaptor_input: AdaptorInput):
        xyz = adaptor_input.adaptor_feature
        if self.mode == "linear":
            out = xyz
        elif self.mode in ("square", "exp"):
            d = xyz.norm(dim=1, keepdim=True)
            out = xyz / d.clip(min=1e-8)
            out = out * (d.square() if self.mode == "square" else torch.expm1(d))
        elif self.mode == "z_exp":
            xy, z = xyz.split([2, 1], dim=1)
            z = torch.exp(z)
            out = torch.cat([xy * z, z], dim=1)
        else:
            raise ValueError(f"Invalid mode: {self.mode}")
        if not self.no_bounds:
            out = out.clip(self.vmin, self.vmax)
        return RegressionAdaptorOutput(value=out)


class DepthAdaptor(UniCeptionAdaptorBase):
    """prediction_heads/adaptors.py:214-257."""

    def __init__(self, name: str, mode: str, vmin: float = 0, vmax: float = math.inf, *args, **kwargs):
        super().__init__(name, required_channels=1)
        self.mode = mode
        self.vmin = vmin
        self.vmax = vmax
        self.no_bounds = (vmin == -float("inf")) and (vmax == float("inf"))

    def forward(self, adaptor_input: AdaptorInput):
        x = adaptor_input.adaptor_feature
        if self.mode == "linear":
            out = x
        elif self.mode == "square":
            out = x ** 2
        elif self.mode == "exp":
            out = torch.exp(x)
        else:
            raise ValueError(f"Invalid mode: {self.mode}")
        if not self.no_bounds:
            out = out.clip(self.vmin, self.vmax)
        return RegressionAdaptorOutput(value=out)


class ConfidenceAdaptor(UniCeptionAdaptorBase):
    """prediction_heads/adaptors.py:1035-1096."""

    def __init__(self, name: str, confidence_type: str, vmin: float, vmax: float, *args, **kwargs):
        super().__init__(name, required_channels=1)
        self.confidence_type = confidence_type
        self.vmin = vmin
        self.vmax = vmax
        assert vmin < vmax, "vmin must be less than vmax"
        if confidence_type == "sigmoid":
            assert math.isfinite(vmin) and math.isfinite(vmax), "vmin and vmax must be finite for sigmoid confidence"
            assert vmin >= 0

    def forward(self, adaptor_input: AdaptorInput):
        x = adaptor_input.adaptor_feature
        if self.confidence_type == "exp":
            return RegressionAdaptorOutput(value=self.vmin + x.exp().clip(max=self.vmax - self.vmin))
        if self.confidence_type == "sigmoid":
            return RegressionAdaptorOutput(value=torch.sigmoid(x) * (self.vmax - self.vmin) + self.vmin)
        if self.confidence_type == "softmax":
            B, C, H, W = x.shape
            return RegressionAdaptorOutput(
                value=torch.nn.functional.softmax(x.reshape(B, C, -1), dim=-1).reshape(B, C, H, W) * (H * W))
        raise ValueError(f"Invalid confidence type: {self.confidence_type}")


class ValueWithConfidenceAdaptor(UniCeptionAdaptorBase):
    """prediction_heads/adaptors.py:1189-1230."""

    def __init__(self, name: str, value_adaptor: UniCeptionAdaptorBase, confidence_adaptor: UniCeptionAdaptorBase, *args, **kwargs):
        super().__init__(name, required_channels=value_adaptor.required_channels + confidence_adaptor.required_channels)
        self.value_adaptor = value_adaptor
        self.confidence_adaptor = confidence_adaptor

    def forward(self, adaptor_input: AdaptorInput):
        v_in, c_in = torch.split(adaptor_input.adaptor_feature,
                                 [self.value_adaptor.required_channels, self.confidence_adaptor.required_channels], dim=1)
        v = self.value_adaptor(AdaptorInput(adaptor_feature=v_in, output_shape_hw=adaptor_input.output_shape_hw))
        c = self.confidence_adaptor(AdaptorInput(adaptor_feature=c_in, output_shape_hw=adaptor_input.output_shape_hw))
        return RegressionWithConfidenceAdaptorOutput(value=v.value, confidence=c.value)


class PointMapWithConfidenceAdaptor(ValueWithConfidenceAdaptor):
    """prediction_heads/adaptors.py:1269-1293."""

    def __init__(self, name: str, pointmap_mode: str, pointmap_vmin: float, pointmap_vmax: float, confidence_type: str,
                 confidence_vmin: float, confidence_vmax: float, *args, **kwargs):
        super().__init__(
            name,
            value_adaptor=PointMapAdaptor(name=f"{name}", mode=pointmap_mode, vmin=pointmap_vmin, vmax=pointmap_vmax),
            confidence_adaptor=ConfidenceAdaptor(name=f"{name}_confidence", confidence_type=confidence_type,
                                                 vmin=confidence_vmin, vmax=confidence_vmax),
        )

    def fusable(self) -> bool:
        """True when the fused head kernel implements exactly this configuration (the DUSt3R default)."""
        v, c = self.value_adaptor, self.confidence_adaptor
        return v.mode == "exp" and v.no_bounds and c.confidence_type == "exp"
