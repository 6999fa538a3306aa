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

from .checkpoints import load_checkpoint_file
from . import dpt_engine, fused, ops
from .autograd_ops import LinearFn
from .params import ParamPack


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
            ckpt = load_checkpoint_file(self.pretrained_checkpoint_path)
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


# ------------------------------------------------------------------------------------------------
# DPT head (prediction_heads/dpt.py:32-311; libs/croco/dpt_block.py)
# ------------------------------------------------------------------------------------------------
@dataclass
class DPTFeatureInput:
    features_upsampled_8x: torch.Tensor
    target_output_shape: Tuple[int, int]


class _ResidualConvUnit(nn.Module):
    """ResidualConvUnit_custom parameters (dpt_block.py:114-177), bn=False."""

    def __init__(self, features: int):
        super().__init__()
        self.conv1 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1, bias=True)
        self.conv2 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1, bias=True)


class _FeatureFusionBlock(nn.Module):
    """FeatureFusionBlock_custom parameters (dpt_block.py:180-255)."""

    def __init__(self, features: int, with_unit1: bool = True):
        super().__init__()
        self.out_conv = nn.Conv2d(features, features, kernel_size=1, stride=1, padding=0, bias=True)
        self.resConfUnit1 = _ResidualConvUnit(features)
        self.resConfUnit2 = _ResidualConvUnit(features)
        if not with_unit1:
            del self.resConfUnit1  # dpt.py:83


class DPTFeature(nn.Module):
    """Parameter container with the reference's keys (incl. the three aliases of every layer_rn conv,
    dpt_block.py:34-78): scratch.layer{1..4}_rn == scratch.layer_rn.{0..3} == input_process.{0..3}.1."""

    def __init__(self, patch_size=16, main_tasks=("rgb",), hooks=(2, 5, 8, 11), input_feature_dims=(768, 768, 768, 768),
                 layer_dims=(96, 192, 384, 768), feature_dim: int = 256, use_bn: bool = False, output_width_ratio=1,
                 pretrained_checkpoint_path: str = None, checkpoint_gradient: bool = False, nonlinearity: str = "relu",
                 *args, **kwargs):
        super().__init__()
        if use_bn or nonlinearity != "relu" or output_width_ratio != 1:
            raise NotImplementedError("uniception_b200.DPTFeature: only use_bn=False, relu, output_width_ratio=1 are built")
        self.patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.main_tasks = main_tasks
        self.hooks = list(hooks)
        self.layer_dims = list(layer_dims)
        self.feature_dim = feature_dim
        self.input_feature_dims = list(input_feature_dims) if not isinstance(input_feature_dims, int) else [input_feature_dims] * 4
        self.checkpoint_gradient = checkpoint_gradient
        assert len(self.hooks) == 4 and len(self.input_feature_dims) == 4 and len(self.layer_dims) == 4
        L = self.layer_dims
        scratch = nn.Module()
        scratch.layer1_rn = nn.Conv2d(L[0], feature_dim, kernel_size=3, stride=1, padding=1, bias=False)
        scratch.layer2_rn = nn.Conv2d(L[1], feature_dim, kernel_size=3, stride=1, padding=1, bias=False)
        scratch.layer3_rn = nn.Conv2d(L[2], feature_dim, kernel_size=3, stride=1, padding=1, bias=False)
        scratch.layer4_rn = nn.Conv2d(L[3], feature_dim, kernel_size=3, stride=1, padding=1, bias=False)
        scratch.layer_rn = nn.ModuleList([scratch.layer1_rn, scratch.layer2_rn, scratch.layer3_rn, scratch.layer4_rn])
        scratch.refinenet1 = _FeatureFusionBlock(feature_dim)
        scratch.refinenet2 = _FeatureFusionBlock(feature_dim)
        scratch.refinenet3 = _FeatureFusionBlock(feature_dim)
        scratch.refinenet4 = _FeatureFusionBlock(feature_dim, with_unit1=False)
        self.scratch = scratch
        D = self.input_feature_dims
        act = [
            nn.Sequential(nn.Conv2d(D[0], L[0], kernel_size=1), nn.ConvTranspose2d(L[0], L[0], kernel_size=4, stride=4, padding=0)),
            nn.Sequential(nn.Conv2d(D[1], L[1], kernel_size=1), nn.ConvTranspose2d(L[1], L[1], kernel_size=2, stride=2, padding=0)),
            nn.Sequential(nn.Conv2d(D[2], L[2], kernel_size=1)),
            nn.Sequential(nn.Conv2d(D[3], L[3], kernel_size=1), nn.Conv2d(L[3], L[3], kernel_size=3, stride=2, padding=1)),
        ]
        self.act_postprocess = act  # plain list (not registered), like the reference
        self.input_process = nn.ModuleList([nn.Sequential(a, l) for a, l in zip(act, scratch.layer_rn)])
        if pretrained_checkpoint_path is not None:
            print(f"Loading pretrained DPT dense feature head from {pretrained_checkpoint_path}")
            ckpt = load_checkpoint_file(pretrained_checkpoint_path)
            print(self.load_state_dict(ckpt["model"]))

    def forward(self, dpt_input: PredictionHeadLayeredInput) -> DPTFeatureInput:
        """prediction_heads/dpt.py:180-232: 4 hooked BCHW maps -> `features_upsampled_8x` [B, feature_dim, 8h', 8w'] (fp32 NCHW
        at the module boundary like the reference; NHWC bf16 inside).  `checkpoint_gradient` only changes what the reference
        keeps for its backward; the 3x3 convs here re-gather their im2col operand in the backward either way."""
        assert self.input_feature_dims is not None, "Need to call init(input_feature_dims) function first"
        layered = dpt_input.list_features
        for hook_idx, hook in enumerate(self.hooks):
            assert layered[hook].shape[1] == self.input_feature_dims[hook_idx], \
                f"Input feature dimension mismatch at hook {hook}. Expected BCHW"
        feats = [layered[hook] for hook in self.hooks]
        if not feats[0].is_cuda:
            raise RuntimeError("uniception_b200.DPTFeature runs on CUDA only (no CPU fallback)")
        B, _, h, w = feats[0].shape
        toks = [fused.NchwToNlcFn.apply(f) for f in feats]
        p1 = _DPTFeatureFn.apply(self, B, h, w, *toks, *list(self.parameters()))
        s0 = self.act_postprocess[0][1].kernel_size[0]  # layer 1 is upsampled by its ConvTranspose, then x2 by refinenet1
        Hf, Wf = 2 * s0 * h, 2 * s0 * w
        fd = self.feature_dim
        x = p1 if p1.shape[1] == fd else p1[:, :fd]
        out = fused.NlcToNchwFn.apply(x.contiguous(), B, Hf, Wf)
        return DPTFeatureInput(features_upsampled_8x=out, target_output_shape=dpt_input.target_output_shape)


class DPTRegressionProcessor(nn.Module):
    """prediction_heads/dpt.py:238-311 parameters: conv1 (3x3, C -> C/2), conv2 = [3x3 C/2 -> hidden, ReLU, 1x1 hidden -> out]."""

    def __init__(self, input_feature_dim: int, output_dim: int, hidden_dims=None, pretrained_checkpoint_path: str = None,
                 checkpoint_gradient: bool = False, *args, **kwargs):
        super().__init__()
        if hidden_dims is None:
            hidden_dims = [input_feature_dim // 2] * 2
        assert isinstance(hidden_dims, (list, tuple)) and len(hidden_dims) == 2
        self.output_dim = output_dim
        self.checkpoint_gradient = checkpoint_gradient
        self.conv1 = nn.Conv2d(input_feature_dim, hidden_dims[0], kernel_size=3, stride=1, padding=1)
        self.conv2 = nn.Sequential(nn.Conv2d(hidden_dims[0], hidden_dims[1], kernel_size=3, stride=1, padding=1), nn.ReLU(True),
                                   nn.Conv2d(hidden_dims[1], output_dim, kernel_size=1, stride=1, padding=0))
        if pretrained_checkpoint_path is not None:
            print(f"Loading pretrained DPT regression processor from {pretrained_checkpoint_path}")
            ckpt = load_checkpoint_file(pretrained_checkpoint_path)
            print(self.load_state_dict(ckpt["model"]))

    def forward(self, dpt_processor_input: DPTFeatureInput) -> PixelTaskOutput:
        """prediction_heads/dpt.py:285-311: `features_upsampled_8x` [B, C, Hf, Wf] -> decoded channels [B, output_dim, H, W]."""
        x = dpt_processor_input.features_upsampled_8x
        H, W = dpt_processor_input.target_output_shape
        if not x.is_cuda:
            raise RuntimeError("uniception_b200.DPTRegressionProcessor runs on CUDA only (no CPU fallback)")
        B, C, Hf, Wf = x.shape
        assert C == self.conv1.weight.shape[1], f"Input feature dimension mismatch: {C} vs {self.conv1.weight.shape[1]}"
        tok = fused.NchwToNlcFn.apply(x)
        cp = dpt_engine._pad64(C)
        if cp != C:  # channel counts are padded to the GEMM granularity (zeros)
            tok = torch.nn.functional.pad(tok, (0, cp - C))
        y = _DPTRegressorFn.apply(self, B, Hf, Wf, (int(H), int(W)), tok, *list(self.parameters()))
        od = self.output_dim
        return PixelTaskOutput(decoded_channels=y[:, :od].reshape(B, H, W, od).permute(0, 3, 1, 2).contiguous())


def _tape_backward(ctx, dy, inputs, weights):
    """Shared backward of the tape-based DPT nodes: seeds the output gradient, replays the tape, moves the GEMM-layout
    weight gradients into the parameters' .grad and returns the input-token gradients."""
    tape = ctx.tape
    g = dy.contiguous()
    tape.add_grad(ctx.y, g if g.dtype == torch.bfloat16 else g.to(torch.bfloat16))
    tape.backward()
    for cw in weights.all():
        cw.flush_grads()
    grads = []
    for t, dt in zip(inputs, ctx.in_dtypes):
        gt = tape.pop_grad(t)
        grads.append(None if gt is None else (gt if gt.dtype == dt else gt.to(dt)))
    return grads


class _DPTFeatureFn(torch.autograd.Function):
    """tokens (4 x bf16 [B*h*w, C_j]) -> features_upsampled_8x, NHWC bf16 [B*Hf*Wf, feature_dim_pad]."""

    @staticmethod
    def forward(ctx, feat_mod, B, h, w, t0, t1, t2, t3, *params):
        toks = [t.contiguous() if t.dtype == torch.bfloat16 else t.to(torch.bfloat16).contiguous() for t in (t0, t1, t2, t3)]
        Wt = dpt_engine.DPTFeatureWeights(feat_mod)
        tape = dpt_engine.Tape()
        y, _, _ = dpt_engine.dpt_feature_forward(tape, Wt, toks, B, h, w)
        ctx.tape, ctx.Wt, ctx.toks, ctx.y = tape, Wt, toks, y
        ctx.in_dtypes = [t.dtype for t in (t0, t1, t2, t3)]
        return y

    @staticmethod
    def backward(ctx, dy):
        grads = _tape_backward(ctx, dy, ctx.toks, ctx.Wt)
        ctx.tape = ctx.Wt = ctx.toks = ctx.y = None
        return (None, None, None, None, *grads) + (None,) * (len(ctx.needs_input_grad) - 8)


class _DPTRegressorFn(torch.autograd.Function):
    """NHWC bf16 [B*Hf*Wf, C_pad] -> raw regression output fp32 [B*H*W, 64] (first output_dim columns valid)."""

    @staticmethod
    def forward(ctx, reg_mod, B, Hf, Wf, out_hw, x, *params):
        xb = x.contiguous() if x.dtype == torch.bfloat16 else x.to(torch.bfloat16).contiguous()
        Wt = dpt_engine.DPTRegressorWeights(reg_mod, xb.shape[1])
        tape = dpt_engine.Tape()
        y = dpt_engine.dpt_regressor_forward(tape, Wt, xb, B, Hf, Wf, out_hw)
        ctx.tape, ctx.Wt, ctx.x, ctx.y = tape, Wt, xb, y
        ctx.in_dtypes = [x.dtype]
        return y

    @staticmethod
    def backward(ctx, dy):
        grads = _tape_backward(ctx, dy, [ctx.x], ctx.Wt)
        ctx.tape = ctx.Wt = ctx.x = ctx.y = None
        return (None, None, None, None, None, grads[0]) + (None,) * (len(ctx.needs_input_grad) - 6)


class _DPTHeadFn(torch.autograd.Function):
    """tokens (4 x bf16 [B*h*w, C_j]) -> raw head output fp32 [B*H*W, 64] (first output_dim columns valid)."""

    @staticmethod
    def forward(ctx, feat_mod, reg_mod, B, h, w, out_hw, t0, t1, t2, t3, *params):
        toks = [t.contiguous() if t.dtype == torch.bfloat16 else t.to(torch.bfloat16).contiguous() for t in (t0, t1, t2, t3)]
        Wt = dpt_engine.DPTWeights(feat_mod, reg_mod)
        tape = dpt_engine.Tape()
        y = dpt_engine.dpt_forward(tape, Wt, toks, B, h, w, out_hw)
        ctx.tape, ctx.Wt, ctx.toks, ctx.y = tape, Wt, toks, y
        ctx.in_dtypes = [t.dtype for t in (t0, t1, t2, t3)]
        return y

    @staticmethod
    def backward(ctx, dy):
        tape, Wt, toks = ctx.tape, ctx.Wt, ctx.toks
        g = dy.contiguous()
        tape.add_grad(ctx.y, g if g.dtype == torch.bfloat16 else g.to(torch.bfloat16))
        tape.backward()
        for cw in Wt.all():
            cw.flush_grads()
        grads = []
        for t, dt in zip(toks, ctx.in_dtypes):
            gt = tape.pop_grad(t)
            grads.append(None if gt is None else (gt if gt.dtype == dt else gt.to(dt)))
        ctx.tape = ctx.Wt = ctx.toks = ctx.y = None
        return (None, None, None, None, None, None, *grads) + (None,) * (len(ctx.needs_input_grad) - 10)


class _HeadPostFn(torch.autograd.Function):
    """pointmap(exp) + confidence(exp) adaptor + BHWC layout on a [B*H*W, ld] fp32 head output (patch = 1)."""

    @staticmethod
    def forward(ctx, y, B, H, W, conf_min, conf_max):
        pts, conf = ops.head_post_fwd(y, B, H, W, 1, conf_min, conf_max)
        ctx.save_for_backward(y)
        ctx.cfg = (B, H, W, conf_min, conf_max)
        return pts, conf

    @staticmethod
    def backward(ctx, dpts, dconf):
        (y,) = ctx.saved_tensors
        B, H, W, cmin, cmax = ctx.cfg
        if dpts is None:
            dpts = torch.zeros(B, H, W, 3, device=y.device)
        if dconf is None:
            dconf = torch.zeros(B, H, W, 1, device=y.device)
        dy = ops.head_post_bwd(y, dpts.float(), dconf.float(), B, H, W, 1, cmin, cmax, dtype=torch.float32)
        return dy, None, None, None, None, None


class DPTHead(nn.Module):
    """`nn.Sequential(DPTFeature, DPTRegressionProcessor)` of factory/dust3r.py:178,192 on the B200 engine.
    Keeps the Sequential's state-dict keys (`0.*`, `1.*`)."""

    def __init__(self, feature: DPTFeature, regressor: DPTRegressionProcessor):
        super().__init__()
        self.add_module("0", feature)
        self.add_module("1", regressor)

    @property
    def feature(self) -> DPTFeature:
        return getattr(self, "0")

    @property
    def regressor(self) -> DPTRegressionProcessor:
        return getattr(self, "1")

    def forward_tokens(self, toks: List[torch.Tensor], B: int, h: int, w: int, out_hw: Tuple[int, int]) -> torch.Tensor:
        params = list(self.parameters())
        return _DPTHeadFn.apply(self.feature, self.regressor, B, h, w, tuple(out_hw), *toks, *params)

    def forward(self, head_input: PredictionHeadLayeredInput) -> PixelTaskOutput:
        feats = head_input.list_features
        assert len(feats) == 4, "DPT head expects 4 hooked feature maps"
        if not feats[0].is_cuda:
            raise RuntimeError("uniception_b200.DPTHead runs on CUDA only (no CPU fallback)")
        B, _, h, w = feats[0].shape
        H, W = head_input.target_output_shape
        toks = [fused.NchwToNlcFn.apply(f) for f in feats]
        y = self.forward_tokens(toks, B, h, w, (H, W))
        od = self.regressor.output_dim
        return PixelTaskOutput(decoded_channels=y[:, :od].reshape(B, H, W, od).permute(0, 3, 1, 2).contiguous())
