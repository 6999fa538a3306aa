"""DPT head (DPTFeature + DPTRegressionProcessor) on the B200 kernels, forward and backward.

Reference arithmetic: prediction_heads/dpt.py:94-232 (feature head), :271-311 (regression processor),
libs/croco/dpt_block.py:114-177 (ResidualConvUnit_custom), :180-255 (FeatureFusionBlock_custom).
Feature maps are NHWC bf16 (token-major [B*H*W, C]); every conv is `uc_gemm` (see csrc/dpt.cu).  The graph
is irregular (4 scales, skips, crops), so instead of a hand-written backward schedule this file keeps a tiny
tape: every forward helper records a closure that consumes the gradient of its output and adds the gradients
of its inputs; `Tape.backward()` replays the closures in reverse.  No torch autograd inside.

Channel counts that are not multiples of 64 (layer_dims[0] = 96, the 4 output channels) are zero-padded to the
next multiple of 64 in the bf16 operand copies, so the GEMM tile constraints hold; padded outputs are exact zeros.
"""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


class Tape:
    def __init__(self):
        self.fns: List[Callable[[], None]] = []
        self.grads: Dict[int, torch.Tensor] = {}
        self.keep: List[torch.Tensor] = []
        # ids of fused-ReLU conv outputs whose single consumer applies the ReLU mask in its own dgrad epilogue
        # (UC_EPI_RELU_BWD): their producers receive the gradient already masked
        self.premasked: set = set()

    def add_grad(self, t: torch.Tensor, g: torch.Tensor) -> None:
        k = id(t)
        if k in self.grads:
            self.grads[k] = ops.elementwise(0, self.grads[k], g)
        else:
            self.grads[k] = g
            self.keep.append(t)

    def pop_grad(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        return self.grads.pop(id(t), None)

    def record(self, fn: Callable[[], None]) -> None:
        self.fns.append(fn)

    def backward(self) -> None:
        for fn in reversed(self.fns):
            fn()
        self.fns.clear()


def _accumulate_grad(p: nn.Parameter, g: torch.Tensor) -> None:
    """p.grad += g.  A parameter that lives in a ParamPack (DUSt3R(pred_head_type="dpt")) accumulates into its view of the
    pack's flat gradient buffer -- after the pack has applied `zero_grad(set_to_none=True)` semantics -- so the value is
    what the optimizer and dp.GradSync see; a free-standing parameter gets a fresh tensor."""
    tag = getattr(p, "_uc_pack", None)
    if tag is not None and tag[0].valid() and tag[0].params.get(tag[1]) is p:
        tag[0].prepare_grads()
    if p.grad is None:
        p.grad = g.contiguous().clone()
    else:
        p.grad.add_(g)


class ConvW:
    """bf16 GEMM operand + fp32 bias of one conv layer, and the fp32 gradient accumulators in GEMM layout."""

    def __init__(self, kind: str, weight: nn.Parameter, bias: Optional[nn.Parameter], cin_pad: int, stride: int = 1):
        self.kind, self.weight, self.bias, self.stride = kind, weight, bias, stride
        w = weight.detach()
        dev = w.device
        if kind == "conv3":  # [co, ci, 3, 3] -> [co_pad, 9 * cin_pad], k index = tap * cin_pad + ci
            co, ci = w.shape[0], w.shape[1]
            self.co, self.ci, self.co_pad, self.ci_pad = co, ci, _pad64(co), cin_pad
            m = torch.zeros(self.co_pad, 3, 3, cin_pad, device=dev)
            m[:co, :, :, :ci] = w.permute(0, 2, 3, 1)
            self.mat = m.view(self.co_pad, 9 * cin_pad)
        elif kind == "conv1":  # [co, ci, 1, 1] -> [co_pad, cin_pad]
            co, ci = w.shape[0], w.shape[1]
            self.co, self.ci, self.co_pad, self.ci_pad = co, ci, _pad64(co), cin_pad
            m = torch.zeros(self.co_pad, cin_pad, device=dev)
            m[:co, :ci] = w.view(co, ci)
            self.mat = m
        elif kind == "convT":  # [ci, co, s, s] -> [(i, j, co_pad), cin_pad]
            ci, co, s = w.shape[0], w.shape[1], w.shape[2]
            self.co, self.ci, self.co_pad, self.ci_pad, self.s = co, ci, _pad64(co), cin_pad, s
            m = torch.zeros(s, s, self.co_pad, cin_pad, device=dev)
            m[:, :, :co, :ci] = w.permute(2, 3, 1, 0)
            self.mat = m.view(s * s * self.co_pad, cin_pad)
        else:
            raise ValueError(kind)
        self.w16 = self.mat.to(torch.bfloat16).contiguous()
        n_out = self.w16.shape[0]
        self.b32 = None
        if bias is not None:
            b = torch.zeros(self.co_pad, device=dev)
            b[:self.co] = bias.detach()
            self.b32 = b.repeat(self.s * self.s) if kind == "convT" else b
        self.gw = None  # fp32 [n_out, K] accumulated by wgrad GEMMs
        self.gb = None
        self.n_out = n_out

    def grad_buffers(self):
        if self.gw is None:
            self.gw = torch.zeros(self.w16.shape, dtype=torch.float32, device=self.w16.device)
            self.gb = torch.zeros(self.n_out, dtype=torch.float32, device=self.w16.device)
        return self.gw, self.gb

    def flush_grads(self) -> None:
        """GEMM-layout fp32 gradients -> the parameters' .grad (layout permutation + un-padding; plumbing)."""
        if self.gw is None:
            return
        co, ci = self.co, self.ci
        if self.kind == "conv3":
            g = self.gw.view(self.co_pad, 3, 3, self.ci_pad)[:co, :, :, :ci].permute(0, 3, 1, 2)
            gb = self.gb[:co]
        elif self.kind == "conv1":
            g = self.gw[:co, :ci].view(co, ci, 1, 1)
            gb = self.gb[:co]
        else:
            s = self.s
            g = self.gw.view(s, s, self.co_pad, self.ci_pad)[:, :, :co, :ci].permute(3, 2, 0, 1)
            gb = self.gb.view(s * s, self.co_pad)[:, :co].sum(0)
        if self.weight.requires_grad:
            _accumulate_grad(self.weight, g)
        if self.bias is not None and self.bias.requires_grad:
            _accumulate_grad(self.bias, gb)
        self.gw = self.gb = None


# ------------------------------------------------------------------------------------------------
# forward helpers (each records its backward on the tape)
# ------------------------------------------------------------------------------------------------
def _gemm_fwd(x, cw: ConvW, out_dtype=torch.bfloat16, relu=False, residual=None):
    out = torch.empty(x.shape[0], cw.n_out, dtype=out_dtype, device=x.device)
    ops.gemm(x, cw.w16, out, bias=cw.b32, relu=relu, residual=residual)
    return out


def _gemm_bwd(tape: Tape, dy, x_in, cw: ConvW, need_dx=True, relu_mask=None):
    """dy [rows, n_out] bf16.  Accumulates weight/bias grads; returns d x_in (times (relu_mask > 0) when given)."""
    gw, gb = cw.grad_buffers()
    ops.gemm(dy, x_in, gw, a_layout=1, b_layout=1, atomic=True)
    if cw.b32 is not None:
        ops.colsum_(dy, gb)
    if not need_dx:
        return None
    dx = torch.empty(dy.shape[0], cw.w16.shape[1], dtype=torch.bfloat16, device=dy.device)
    if relu_mask is not None:
        ops.gemm(dy, cw.w16, dx, b_layout=1, relu_bwd=True, aux_in=relu_mask)
    else:
        ops.gemm(dy, cw.w16, dx, b_layout=1)
    return dx


def conv1x1(tape: Tape, x, cw: ConvW, out_dtype=torch.bfloat16, need_dx=True, x_premask=False):
    """x_premask: x is the output of a fused-ReLU conv and this is its only consumer -- the ReLU mask is applied by this
    conv's dgrad epilogue instead of a separate pass in the producer's backward."""
    y = _gemm_fwd(x, cw, out_dtype)
    if x_premask:
        tape.premasked.add(id(x))

    def bwd():
        g = tape.pop_grad(y)
        if g is None:
            return
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        dx = _gemm_bwd(tape, g.contiguous(), x, cw, need_dx, relu_mask=x if x_premask else None)
        if dx is not None:
            tape.add_grad(x, dx)

    tape.record(bwd)
    return y


IMPLICIT_CONV = os.environ.get("UC_CONV_IM2COL", "0") != "1"  # UC_CONV_IM2COL=1: the materialised-column path (A/B, debugging)


def conv3x3(tape: Tape, x, B, H, W, cw: ConvW, relu=False, residual=None, x_relu_src=None, x_premask=False):
    """3x3, pad 1, stride cw.stride.  Optional fused ReLU or fused `+ residual` (one of the two).
    Stride 1: implicit GEMM (uc_conv3x3: shifted TMA boxes, no column buffer) forward, dgrad and wgrad.
    Stride 2 (one small layer of the head, dpt.py:118-126): uc_im2col3x3 -> uc_gemm.
    ReLU backward rides the dgrad epilogue (UC_EPI_RELU_BWD) when x is a ReLU output with this conv as its ONLY consumer:
    x_relu_src = z with x = relu(z) from `relu()`: the masked dx is delivered to z directly; x_premask: x came out of a conv
    with a fused ReLU, which then receives its gradient already masked."""
    st = cw.stride
    if st == 1 and IMPLICIT_CONV:
        y = ops.conv3x3_fwd(x, cw.w16, B, H, W, bias=cw.b32, relu=relu, residual=residual)
        if x_premask:
            tape.premasked.add(id(x))

        def bwd_implicit():
            g = tape.pop_grad(y)
            if g is None:
                return
            if residual is not None:
                tape.add_grad(residual, g)
            if relu and id(y) not in tape.premasked:
                g = ops.elementwise(2, g, y)  # g * (y > 0)
            gw, gb = cw.grad_buffers()
            ops.conv3x3_wgrad_(x, g, gw, B, H, W)
            if cw.b32 is not None:
                ops.colsum_(g, gb)
            if x_relu_src is not None:
                tape.add_grad(x_relu_src, ops.conv3x3_dgrad(g, cw.w16, B, H, W, relu_out=x))
            elif x_premask:
                tape.add_grad(x, ops.conv3x3_dgrad(g, cw.w16, B, H, W, relu_out=x))
            else:
                tape.add_grad(x, ops.conv3x3_dgrad(g, cw.w16, B, H, W))

        tape.record(bwd_implicit)
        return y
    cols = ops.im2col3x3(x, B, H, W, st)
    y = _gemm_fwd(cols, cw, relu=relu, residual=residual)
    del cols  # re-gathered in backward: trades one memory-bound pass for not holding 9x the activation

    def bwd():
        g = tape.pop_grad(y)
        if g is None:
            return
        if residual is not None:
            tape.add_grad(residual, g)
        if relu:
            g = ops.elementwise(2, g, y)  # g * (y > 0)
        cols_b = ops.im2col3x3(x, B, H, W, st)
        dcols = _gemm_bwd(tape, g, cols_b, cw)
        del cols_b
        tape.add_grad(x, ops.col2im3x3(dcols, B, H, W, st))

    tape.record(bwd)
    return y


def conv_transpose(tape: Tape, x, B, h, w, cw: ConvW):
    """ConvTranspose2d with kernel == stride == s: GEMM to (i, j, co) columns, then depth-to-space."""
    s = cw.s
    deep = _gemm_fwd(x, cw)
    y = ops.depth_to_space(deep, B, h, w, s)
    del deep

    def bwd():
        g = tape.pop_grad(y)
        if g is None:
            return
        gdeep = ops.space_to_depth(g, B, h, w, s)
        tape.add_grad(x, _gemm_bwd(tape, gdeep, x, cw))

    tape.record(bwd)
    return y


def relu(tape: Tape, x):
    y = ops.elementwise(1, x)

    def bwd():
        g = tape.pop_grad(y)
        if g is not None:
            tape.add_grad(x, ops.elementwise(2, g, y))

    tape.record(bwd)
    return y


def add(tape: Tape, a, b):
    y = ops.elementwise(0, a, b)

    def bwd():
        g = tape.pop_grad(y)
        if g is not None:
            tape.add_grad(a, g)
            tape.add_grad(b, g)

    tape.record(bwd)
    return y


def resize(tape: Tape, x, B, Hi, Wi, Ho, Wo):
    y = ops.bilinear_fwd(x, B, Hi, Wi, Ho, Wo)

    def bwd():
        g = tape.pop_grad(y)
        if g is not None:
            tape.add_grad(x, ops.bilinear_bwd(g, B, Hi, Wi, Ho, Wo))

    tape.record(bwd)
    return y


def crop(tape: Tape, x, B, H, W, Hc, Wc):
    """x[:, :Hc, :Wc] on an NHWC map (index op; torch slicing as plumbing)."""
    if Hc == H and Wc == W:
        return x
    C = x.shape[1]
    y = x.view(B, H, W, C)[:, :Hc, :Wc].reshape(B * Hc * Wc, C).contiguous()

    def bwd():
        g = tape.pop_grad(y)
        if g is not None:
            full = torch.zeros(B, H, W, C, dtype=g.dtype, device=g.device)
            full[:, :Hc, :Wc] = g.view(B, Hc, Wc, C)
            tape.add_grad(x, full.view(B * H * W, C))

    tape.record(bwd)
    return y


# ------------------------------------------------------------------------------------------------
# the head
# ------------------------------------------------------------------------------------------------
class DPTFeatureWeights:
    """Operand copies of one DPTFeature (prediction_heads/dpt.py:94-177), rebuilt per forward from the fp32 masters."""

    def __init__(self, feat: nn.Module):
        sc = feat.scratch
        self.pre, self.up, self.rn = [], [], []
        for j in range(4):
            act = feat.act_postprocess[j]
            c1 = ConvW("conv1", act[0].weight, act[0].bias, cin_pad=act[0].weight.shape[1])
            self.pre.append(c1)
            if j == 0 or j == 1:
                self.up.append(ConvW("convT", act[1].weight, act[1].bias, cin_pad=c1.co_pad))
            elif j == 3:
                self.up.append(ConvW("conv3", act[1].weight, act[1].bias, cin_pad=c1.co_pad, stride=2))
            else:
                self.up.append(None)
            cin = self.up[j].co_pad if self.up[j] is not None else c1.co_pad
            self.rn.append(ConvW("conv3", sc.layer_rn[j].weight, None, cin_pad=cin))
        f = self.rn[0].co_pad
        self.feature_dim_pad = f
        self.fuse = []
        for k in (1, 2, 3, 4):
            blk = getattr(sc, f"refinenet{k}")
            d = {"out": ConvW("conv1", blk.out_conv.weight, blk.out_conv.bias, cin_pad=f)}
            for u in ("resConfUnit1", "resConfUnit2"):
                if hasattr(blk, u):
                    unit = getattr(blk, u)
                    d[u] = (ConvW("conv3", unit.conv1.weight, unit.conv1.bias, cin_pad=f),
                            ConvW("conv3", unit.conv2.weight, unit.conv2.bias, cin_pad=f))
            self.fuse.append(d)

    def all(self):
        out = self.pre + [u for u in self.up if u is not None] + self.rn
        for d in self.fuse:
            out.append(d["out"])
            for u in ("resConfUnit1", "resConfUnit2"):
                if u in d:
                    out += list(d[u])
        return out


class DPTRegressorWeights:
    """Operand copies of one DPTRegressionProcessor (prediction_heads/dpt.py:271-283); cin_pad = padded input channels."""

    def __init__(self, reg: nn.Module, cin_pad: int):
        self.r1 = ConvW("conv3", reg.conv1.weight, reg.conv1.bias, cin_pad=cin_pad)
        self.r2 = ConvW("conv3", reg.conv2[0].weight, reg.conv2[0].bias, cin_pad=self.r1.co_pad)
        self.r3 = ConvW("conv1", reg.conv2[2].weight, reg.conv2[2].bias, cin_pad=self.r2.co_pad)

    def all(self):
        return [self.r1, self.r2, self.r3]


class DPTWeights:
    """Operand copies of one (DPTFeature, DPTRegressionProcessor) pair."""

    def __init__(self, feat: nn.Module, reg: nn.Module):
        self.feat = DPTFeatureWeights(feat)
        self.reg = DPTRegressorWeights(reg, self.feat.feature_dim_pad)

    def all(self):
        return self.feat.all() + self.reg.all()


def _rcu(tape, z, B, H, W, unit, extra_skip=None):
    """z + conv2(relu(conv1(relu(z)))) [+ extra_skip]; ReLU of conv1 fused in its epilogue, skip fused in conv2's."""
    skip = z if extra_skip is None else add(tape, z, extra_skip)
    fused = IMPLICIT_CONV and unit[0].stride == 1 and unit[1].stride == 1
    t = conv3x3(tape, relu(tape, z), B, H, W, unit[0], relu=True, x_relu_src=z if fused else None)
    return conv3x3(tape, t, B, H, W, unit[1], residual=skip, x_premask=fused)


def _fusion(tape, a, b, B, H, W, d):
    """FeatureFusionBlock_custom: a [+ RCU1(b)] -> RCU2 -> bilinear x2 (align_corners) -> 1x1 conv."""
    out = a if b is None else _rcu(tape, b, B, H, W, d["resConfUnit1"], extra_skip=a)
    out = _rcu(tape, out, B, H, W, d["resConfUnit2"])
    # out_conv(interpolate(x)) == interpolate(out_conv(x)): a 1x1 conv is per-pixel affine and the bilinear weights sum to 1
    # (dpt_block.py:251-255).  The conv runs on the low-resolution map (a quarter of the FLOPs) with fp32 output and the
    # resampling reads that, so the pair costs ONE bf16 rounding instead of two.
    out = conv1x1(tape, out, d["out"], out_dtype=torch.float32)
    return resize(tape, out, B, H, W, 2 * H, 2 * W)


def dpt_feature_forward(tape: Tape, Wt: DPTFeatureWeights, feats: List[torch.Tensor], B: int, h: int, w: int):
    """DPTFeature.forward (prediction_heads/dpt.py:180-232).  feats: 4 token tensors bf16 [B*h*w, C_j] (already selected by
    `hooks`).  Returns (features_upsampled_8x as NHWC bf16 [B*Hf*Wf, feature_dim_pad], Hf, Wf)."""
    maps, sizes = [], []
    for j in range(4):
        a = conv1x1(tape, feats[j], Wt.pre[j])
        if j == 0 or j == 1:
            s = Wt.up[j].s
            x, hw = conv_transpose(tape, a, B, h, w, Wt.up[j]), (h * s, w * s)
        elif j == 2:
            x, hw = a, (h, w)
        else:
            x, hw = conv3x3(tape, a, B, h, w, Wt.up[j]), ops.conv_out_hw(h, w, 2)
        maps.append(conv3x3(tape, x, B, hw[0], hw[1], Wt.rn[j]))
        sizes.append(hw)
    l0, l1, l2, l3 = maps
    p4 = _fusion(tape, l3, None, B, sizes[3][0], sizes[3][1], Wt.fuse[3])
    p4 = crop(tape, p4, B, 2 * sizes[3][0], 2 * sizes[3][1], sizes[2][0], sizes[2][1])  # dpt.py:213
    p3 = _fusion(tape, p4, l2, B, sizes[2][0], sizes[2][1], Wt.fuse[2])
    p2 = _fusion(tape, p3, l1, B, sizes[1][0], sizes[1][1], Wt.fuse[1])
    p1 = _fusion(tape, p2, l0, B, sizes[0][0], sizes[0][1], Wt.fuse[0])
    return p1, 2 * sizes[0][0], 2 * sizes[0][1]


def dpt_regressor_forward(tape: Tape, Wt: DPTRegressorWeights, p1: torch.Tensor, B: int, Hf: int, Wf: int, out_hw: Tuple[int, int]):
    """DPTRegressionProcessor.forward (prediction_heads/dpt.py:285-311): 3x3 conv -> bilinear (align_corners) to the exact
    target size -> 3x3 conv + ReLU -> 1x1.  Returns y fp32 [B*H*W, 64] (first `output_dim` columns valid)."""
    c1 = conv3x3(tape, p1, B, Hf, Wf, Wt.r1)
    u = resize(tape, c1, B, Hf, Wf, out_hw[0], out_hw[1])
    c2 = conv3x3(tape, u, B, out_hw[0], out_hw[1], Wt.r2, relu=True)
    return conv1x1(tape, c2, Wt.r3, out_dtype=torch.float32, x_premask=IMPLICIT_CONV)


def dpt_forward(tape: Tape, Wt: DPTWeights, feats: List[torch.Tensor], B: int, h: int, w: int, out_hw: Tuple[int, int]):
    """Feature head + regression processor on one tape.  Returns y fp32 [B*H*W, 64] (first `out_dim` columns valid)."""
    p1, Hf, Wf = dpt_feature_forward(tape, Wt.feat, feats, B, h, w)
    return dpt_regressor_forward(tape, Wt.reg, p1, B, Hf, Wf, out_hw)
