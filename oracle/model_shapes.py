"""TEST INFRASTRUCTURE ONLY -- state-dict key -> shape tables of the reference models on the hot path, written out from
the reference's constructors so that the oracle (`dust3r_oracle.py`) can be given weights WITHOUT importing either the
reference or the product package (bench.py's `--impl reference` port fallback must not load `libuc_b200.so`).

Checked against the real reference in tests/test_oracle_golden.py (skipped where /root/reference is absent).
Only the keys the oracle reads are listed (the DPT state dict aliases some tensors under several names,
libs/croco/dpt_block.py:34-78; the oracle reads `input_process.*` and `scratch.layer_rn.*`).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

Shapes = Dict[str, Tuple[int, ...]]


def _ln(out: Shapes, name: str, c: int):
    out[name + ".weight"] = (c,)
    out[name + ".bias"] = (c,)


def _lin(out: Shapes, name: str, o: int, i: int):
    out[name + ".weight"] = (o, i)
    out[name + ".bias"] = (o,)


def encoder_shapes(prefix="encoder.", C=1024, depth=24, patch=16, mlp_ratio=4) -> Shapes:
    """CroCoEncoder (encoders/croco.py:61-97; libs/croco/blocks.py:64-161)."""
    s: Shapes = {}
    s[prefix + "patch_embed.proj.weight"] = (C, 3, patch, patch)
    s[prefix + "patch_embed.proj.bias"] = (C,)
    for i in range(depth):
        b = f"{prefix}enc_blocks.{i}."
        _ln(s, b + "norm1", C)
        _lin(s, b + "attn.qkv", 3 * C, C)
        _lin(s, b + "attn.proj", C, C)
        _ln(s, b + "norm2", C)
        _lin(s, b + "mlp.fc1", mlp_ratio * C, C)
        _lin(s, b + "mlp.fc2", C, mlp_ratio * C)
    _ln(s, prefix + "enc_norm", C)
    return s


def decoder_shapes(prefix="info_sharing.", C_in=1024, C=768, depth=12, views=2, mlp_ratio=4) -> Shapes:
    """MultiViewCrossAttentionTransformer (info_sharing/cross_attention_transformer.py:97-150;
    utils/transformer_blocks.py:517-647)."""
    s: Shapes = {}
    _lin(s, prefix + "proj_embed", C, C_in)
    for v in range(views):
        for i in range(depth):
            b = f"{prefix}multi_view_branches.{v}.{i}."
            for n in ("norm1", "norm_y", "norm2", "norm3"):
                _ln(s, b + n, C)
            _lin(s, b + "attn.qkv", 3 * C, C)
            _lin(s, b + "attn.proj", C, C)
            for n in ("projq", "projk", "projv", "proj"):
                _lin(s, b + "cross_attn." + n, C, C)
            _lin(s, b + "mlp.fc1", mlp_ratio * C, C)
            _lin(s, b + "mlp.fc2", C, mlp_ratio * C)
    _ln(s, prefix + "norm", C)
    return s


def dpt_shapes(feat_prefix: str, reg_prefix: str, in_dims, out_dim: int, layer_dims=(96, 192, 384, 768), f=256) -> Shapes:
    """DPTFeature + DPTRegressionProcessor (prediction_heads/dpt.py:94-177, :271-283; libs/croco/dpt_block.py)."""
    s: Shapes = {}
    L = list(layer_dims)
    for j in range(4):
        q = f"{feat_prefix}input_process.{j}.0."
        s[q + "0.weight"] = (L[j], in_dims[j], 1, 1)
        s[q + "0.bias"] = (L[j],)
        if j == 0:
            s[q + "1.weight"] = (L[j], L[j], 4, 4)
            s[q + "1.bias"] = (L[j],)
        elif j == 1:
            s[q + "1.weight"] = (L[j], L[j], 2, 2)
            s[q + "1.bias"] = (L[j],)
        elif j == 3:
            s[q + "1.weight"] = (L[j], L[j], 3, 3)
            s[q + "1.bias"] = (L[j],)
        s[f"{feat_prefix}scratch.layer_rn.{j}.weight"] = (f, L[j], 3, 3)
    for k in (1, 2, 3, 4):
        r = f"{feat_prefix}scratch.refinenet{k}."
        s[r + "out_conv.weight"] = (f, f, 1, 1)
        s[r + "out_conv.bias"] = (f,)
        for u in ("resConfUnit1", "resConfUnit2"):
            if k == 4 and u == "resConfUnit1":  # dpt.py:83
                continue
            for c in ("conv1", "conv2"):
                s[f"{r}{u}.{c}.weight"] = (f, f, 3, 3)
                s[f"{r}{u}.{c}.bias"] = (f,)
    s[reg_prefix + "conv1.weight"] = (f // 2, f, 3, 3)
    s[reg_prefix + "conv1.bias"] = (f // 2,)
    s[reg_prefix + "conv2.0.weight"] = (f // 2, f // 2, 3, 3)
    s[reg_prefix + "conv2.0.bias"] = (f // 2,)
    s[reg_prefix + "conv2.2.weight"] = (out_dim, f // 2, 1, 1)
    s[reg_prefix + "conv2.2.bias"] = (out_dim,)
    return s


def dust3r_shapes(head: str = "linear", patch: int = 16) -> Shapes:
    """factory/dust3r.py:111-203 with its hard-coded ViT-L encoder / base decoder."""
    s = encoder_shapes(patch=patch)
    s.update(decoder_shapes())
    for k in (1, 2):
        if head == "linear":
            s[f"head{k}.linear.weight"] = (4 * patch * patch, 768, 1, 1)
            s[f"head{k}.linear.bias"] = (4 * patch * patch,)
        else:
            s.update(dpt_shapes(f"dpt_feature_head{k}.", f"dpt_regressor_head{k}.", [1024, 768, 768, 768], 4))
    return s


def c5_shapes(patch: int = 14) -> Shapes:
    """BASELINE configs[4]: ViT-L/14 intermediate-feature encoder + DPT depth head."""
    s = encoder_shapes(patch=patch)
    s.update(dpt_shapes("dpt_feature_head.", "dpt_regressor_head.", [1024] * 4, 1))
    return s


def random_init(shapes: Shapes, seed: int = 42) -> Dict[str, torch.Tensor]:
    """Reference-style initialisation without the reference: xavier-uniform matrices, zero Linear biases, LayerNorm 1 / 0
    (encoders/croco.py:129-145), kaiming-uniform-like convs for the heads (torch defaults).  Values differ from the
    reference's RNG stream; the timing baselines only need the architecture and sane magnitudes."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if len(shp) >= 2:
            fan_out, fan_in = shp[0], int(math.prod(shp[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif "norm" in k and k.endswith("weight"):
            sd[k] = torch.ones(shp)
        else:
            sd[k] = torch.zeros(shp)
    return sd
