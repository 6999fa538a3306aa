"""Whole-module autograd Functions: one node per encoder / decoder / head, hand-scheduled inside.

Parameter tensors are passed as (ignored) inputs only so that autograd schedules the node; their
gradients are accumulated in place into the ParamPack's flat fp32 buffer (`param.grad` views) by the
backward kernels, and `None` is returned for them.  This is the "accumulate into main_grad" scheme
used for fused wgrad in large-scale trainers; it is what lets dp.py all-reduce one flat buffer.
"""
from __future__ import annotations

import torch

from . import engine as E
from . import ops
from .params import ParamPack


def _prep_grads(pk: ParamPack) -> None:
    """`zero_grad(set_to_none=True)` semantics: a dropped .grad means zero (see ParamPack.prepare_grads)."""
    pk.prepare_grads()


class EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, pk: ParamPack, prefix: str, cfg: dict, *params):
        y, inter, saved = E.encoder_fwd(pk, prefix, img.contiguous().float(), cfg["depth"], cfg["heads"], cfg["patch"],
                                        cfg["rope_base"], cfg["rope_f0"], cfg.get("take", ()), cfg.get("norm_intermediate", True),
                                        cfg.get("portrait"))
        ctx.pk, ctx.prefix, ctx.cfg, ctx.saved = pk, prefix, cfg, saved
        ctx.n_inter = len(inter)
        return (y, *inter)

    @staticmethod
    def backward(ctx, dy, *dinter):
        pk = ctx.pk
        _prep_grads(pk)
        E.encoder_bwd(pk, ctx.prefix, ctx.saved, dy, ctx.cfg["depth"], ctx.cfg["heads"], dinter)
        ctx.saved = None
        return (None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 4)


class DecoderFn(torch.autograd.Function):
    """inputs: per-view token tensors (bf16 [B*N, C_in]); outputs: per-view final tokens, then the
    flattened intermediates (level-major)."""

    @staticmethod
    def forward(ctx, pk: ParamPack, prefix: str, cfg: dict, nv: int, *tensors):
        toks = [t.contiguous() if t.dtype == torch.bfloat16 else t.to(torch.bfloat16).contiguous() for t in tensors[:nv]]
        outs, inter, saved = E.decoder_fwd(pk, prefix, toks, cfg["B"], cfg["h"], cfg["w"], cfg["depth"], cfg["heads"],
                                           cfg["rope_base"], cfg["rope_f0"], cfg.get("take", ()), cfg.get("norm_intermediate", True),
                                           cfg["has_proj_embed"], cfg["has_norm_y"], cfg.get("softmax_scaling"))
        ctx.pk, ctx.prefix, ctx.cfg, ctx.saved, ctx.nv = pk, prefix, cfg, saved, nv
        ctx.n_levels = len(inter)
        ctx.in_dtypes = [t.dtype for t in tensors[:nv]]
        flat = [t for lvl in inter for t in lvl]
        return (*outs, *flat)

    @staticmethod
    def backward(ctx, *grads):
        pk, nv = ctx.pk, ctx.nv
        _prep_grads(pk)
        d_outs = list(grads[:nv])
        d_inter = [list(grads[nv + l * nv: nv + (l + 1) * nv]) for l in range(ctx.n_levels)]
        need_in = any(ctx.needs_input_grad[4:4 + nv])
        d_in = E.decoder_bwd(pk, ctx.prefix, ctx.saved, d_outs, ctx.cfg["depth"], ctx.cfg["heads"], d_inter,
                             ctx.cfg["has_proj_embed"], ctx.cfg["has_norm_y"], need_input_grad=need_in)
        ctx.saved = None
        d_in = [g if (g is None or g.dtype == dt) else g.to(dt) for g, dt in zip(d_in, ctx.in_dtypes)]
        return (None, None, None, None, *d_in) + (None,) * (len(ctx.needs_input_grad) - 4 - nv)


class MultiViewSelfAttnFn(torch.autograd.Function):
    """Global / alternating self-attention info sharing on the assembled token sequence: x_in [B*L, C_in] (rows: batch,
    [view, token], global extras) -> final normalised tokens [B*L, dim], then one [B*L, dim] tensor per tapped depth."""

    @staticmethod
    def forward(ctx, pk: ParamPack, prefix: str, cfg: dict, x_in, *params):
        xb = x_in.contiguous() if x_in.dtype == torch.bfloat16 else x_in.to(torch.bfloat16).contiguous()
        y, inter, saved = E.mv_self_attn_fwd(pk, prefix, xb, cfg["B"], cfg["nv"], cfg["n_view"], cfg["n_extra"], cfg["h"], cfg["w"],
                                             cfg["depth"], cfg["heads"], cfg["rope_base"], cfg["rope_f0"], cfg["alternating"],
                                             cfg["view_pe"], cfg["has_proj_embed"], cfg.get("softmax_scaling"), cfg.get("take", ()),
                                             cfg.get("norm_intermediate", True), cfg.get("recompute", False))
        ctx.pk, ctx.prefix, ctx.cfg, ctx.saved = pk, prefix, cfg, saved
        ctx.in_dtype = x_in.dtype
        return (y, *inter)

    @staticmethod
    def backward(ctx, *grads):
        pk, cfg = ctx.pk, ctx.cfg
        _prep_grads(pk)
        d_in = E.mv_self_attn_bwd(pk, ctx.prefix, ctx.saved, grads[0], cfg["depth"], cfg["heads"], cfg["alternating"],
                                  cfg["has_proj_embed"], need_input_grad=ctx.needs_input_grad[3], d_inter=list(grads[1:]))
        ctx.saved = None
        if d_in is not None and d_in.dtype != ctx.in_dtype:
            d_in = d_in.to(ctx.in_dtype)
        return (None, None, None, d_in) + (None,) * (len(ctx.needs_input_grad) - 4)


class LinearHeadFn(torch.autograd.Function):
    """LinearFeature + pixel_shuffle + PointMapWithConfidenceAdaptor(exp, exp) + BHWC permute, fused."""

    @staticmethod
    def forward(ctx, tok, pk: ParamPack, prefix: str, cfg: dict, *params):
        t = tok.contiguous() if tok.dtype == torch.bfloat16 else tok.to(torch.bfloat16).contiguous()
        pts, conf, saved = E.linear_head_fwd(pk, prefix, t, cfg["B"], cfg["h"], cfg["w"], cfg["patch"], cfg["conf_min"], cfg["conf_max"])
        ctx.pk, ctx.prefix, ctx.cfg, ctx.saved = pk, prefix, cfg, saved
        ctx.in_dtype = tok.dtype
        return pts, conf

    @staticmethod
    def backward(ctx, dpts, dconf):
        pk, cfg = ctx.pk, ctx.cfg
        _prep_grads(pk)
        if dpts is None:
            dpts = torch.zeros(cfg["B"], cfg["h"] * cfg["patch"], cfg["w"] * cfg["patch"], 3, device=ctx.saved[1].device)
        if dconf is None:
            dconf = torch.zeros(cfg["B"], cfg["h"] * cfg["patch"], cfg["w"] * cfg["patch"], 1, device=ctx.saved[1].device)
        dx = E.linear_head_bwd(pk, ctx.prefix, ctx.saved, dpts.float(), dconf.float(), cfg["B"], cfg["h"], cfg["w"], cfg["patch"],
                               cfg["conf_min"], cfg["conf_max"], need_dx=ctx.needs_input_grad[0])
        ctx.saved = None
        if dx is not None and dx.dtype != ctx.in_dtype:
            dx = dx.to(ctx.in_dtype)
        return (dx, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 4)


class NlcToNchwFn(torch.autograd.Function):
    """tokens [B*L, C] (bf16) -> fp32 [B, C, h, w]; backward is the inverse layout kernel."""

    @staticmethod
    def forward(ctx, tok, B, h, w):
        ctx.dims = (B, h, w, tok.dtype)
        return ops.nlc_to_nchw(tok.view(B, h * w, -1), h, w)

    @staticmethod
    def backward(ctx, g):
        B, h, w, dt = ctx.dims
        out = ops.nchw_to_nlc(g.contiguous().float(), torch.bfloat16 if dt == torch.bfloat16 else torch.float32)
        return out.view(B * h * w, -1), None, None, None


class NchwToNlcFn(torch.autograd.Function):
    """fp32 [B, C, h, w] -> bf16 tokens [B*h*w, C]."""

    @staticmethod
    def forward(ctx, x):
        B, C, h, w = x.shape
        ctx.dims = (B, h, w)
        return ops.nchw_to_nlc(x.contiguous().float(), torch.bfloat16).view(B * h * w, C)

    @staticmethod
    def backward(ctx, g):
        B, h, w = ctx.dims
        return ops.nlc_to_nchw(g.contiguous().view(B, h * w, -1), h, w)
