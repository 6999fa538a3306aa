"""Checkpoint compatibility: original CroCo / DUSt3R / MASt3R checkpoints -> UniCeption-format
state dicts, which `uniception_b200` modules load unchanged (identical keys / shapes).

Restates the key maps of the reference's `examples/models/dust3r/convert_dust3r_weights_to_uniception.py`:
  * decoder            :36-48   decoder_embed -> proj_embed, dec_blocks -> multi_view_branches.0, dec_blocks2 ->
                                multi_view_branches.1 (duplicated from dec_blocks when absent, :28-34), dec_norm -> norm
  * DPT heads          :70-104  downstream_head{k}.dpt.* -> DPTFeature keys; downstream_head{k}.dpt.head.{0,2,4} ->
                                DPTRegressionProcessor conv1 / conv2.0 / conv2.2
  * linear heads       :123-148 downstream_head{k}.proj (nn.Linear [1024,768]) -> LinearFeature.linear (1x1 conv)
  * MASt3R DPT heads   :151-212 as DUSt3R's, ignoring `head_local_features`
  * encoder            encoders/croco.py:99-111: the CroCo encoder keys (patch_embed / enc_blocks / enc_norm) are used as is
Pure dictionary transforms on CPU tensors (no arithmetic): every output tensor is the input tensor (or a view of it).
"""
from __future__ import annotations

from typing import Dict, Mapping

import torch

SD = Dict[str, torch.Tensor]

_DPT_PROCESSOR_KEYS = {"0.weight": "conv1.weight", "0.bias": "conv1.bias", "2.weight": "conv2.0.weight",
                       "2.bias": "conv2.0.bias", "4.weight": "conv2.2.weight", "4.bias": "conv2.2.bias"}


def _model(ckpt: Mapping) -> Mapping[str, torch.Tensor]:
    return ckpt["model"] if "model" in ckpt and isinstance(ckpt["model"], Mapping) else ckpt


def encoder_state_dict(ckpt: Mapping) -> SD:
    """CroCo-v2 / DUSt3R encoder tensors under the names `CroCoEncoder` uses (no renaming: croco.py:99-111)."""
    sd = _model(ckpt)
    return {k: v for k, v in sd.items() if k.startswith(("patch_embed.", "enc_blocks.", "enc_norm."))}


def cross_attention_state_dict(ckpt: Mapping) -> SD:
    """`MultiViewCrossAttentionTransformer(IFR)` state dict from the original decoder tensors (convert script :20-48)."""
    sd = {k: v for k, v in _model(ckpt).items() if "dec" in k}
    if not any(k.startswith("dec_blocks2") for k in sd):  # CroCo checkpoints: one decoder shared by both views
        sd.update({k.replace("dec_blocks", "dec_blocks2"): v for k, v in list(sd.items()) if k.startswith("dec_blocks")})
    out: SD = {}
    for k, v in sd.items():
        if "decoder_embed" in k:
            out[k.replace("decoder_embed", "proj_embed")] = v
        elif "dec_blocks." in k:
            out[k.replace("dec_blocks.", "multi_view_branches.0.")] = v
        elif "dec_blocks2." in k:
            out[k.replace("dec_blocks2.", "multi_view_branches.1.")] = v
        elif "dec_norm" in k:
            out[k.replace("dec_norm", "norm")] = v
    return out


def linear_head_state_dict(ckpt: Mapping, head: int) -> SD:
    """`LinearFeature` state dict of downstream_head{head}: the nn.Linear becomes the 1x1 conv (convert script :123-148)."""
    pre = f"downstream_head{head}.proj."
    p = {k[len(pre):]: v for k, v in _model(ckpt).items() if k.startswith(pre)}
    assert set(p) == {"weight", "bias"}, f"unexpected linear-head keys {sorted(p)}"
    w = p["weight"]
    return {"linear.weight": w.reshape(w.shape[0], w.shape[1], 1, 1), "linear.bias": p["bias"]}


def dpt_head_state_dicts(ckpt: Mapping, head: int):
    """(DPTFeature state dict, DPTRegressionProcessor state dict) of downstream_head{head} (convert script :70-104,
    :151-212; MASt3R's `head_local_features` branch is ignored as there)."""
    pre = f"downstream_head{head}."
    h = {k[len(pre):]: v for k, v in _model(ckpt).items() if k.startswith(pre)}
    h = {(k[len("dpt."):] if k.startswith("dpt.") else k): v for k, v in h.items()}
    feature = {k: v for k, v in h.items() if not k.startswith("head")}
    proc = {k[len("head."):]: v for k, v in h.items() if k.startswith("head.") }
    proc = {_DPT_PROCESSOR_KEYS.get(k, k): v for k, v in proc.items()}
    return feature, proc


def dust3r_state_dict(ckpt: Mapping, pred_head_type: str = "linear") -> SD:
    """Full `uniception_b200.DUSt3R` state dict (encoder.* / info_sharing.* / head{1,2}.* and, for DPT, the
    dpt_feature_head{k}.* / dpt_regressor_head{k}.* aliases the reference registers, factory/dust3r.py:164-192)."""
    out: SD = {"encoder." + k: v for k, v in encoder_state_dict(ckpt).items()}
    out.update({"info_sharing." + k: v for k, v in cross_attention_state_dict(ckpt).items()})
    for h in (1, 2):
        if pred_head_type == "linear":
            out.update({f"head{h}." + k: v for k, v in linear_head_state_dict(ckpt, h).items()})
        elif pred_head_type == "dpt":
            feat, proc = dpt_head_state_dicts(ckpt, h)
            for k, v in feat.items():
                out[f"dpt_feature_head{h}." + k] = v
                out[f"head{h}.0." + k] = v
            for k, v in proc.items():
                out[f"dpt_regressor_head{h}." + k] = v
                out[f"head{h}.1." + k] = v
        else:
            raise ValueError(f"Invalid prediction head type: {pred_head_type}. Must be 'linear' or 'dpt'.")
    return out


def load_checkpoint_file(path: str):
    """`torch.load` for UniCeption checkpoint files ({"model": state_dict, plus plain str / number metadata}).  Tensors-only
    unpickling first (a checkpoint file from an untrusted source must not execute code); files that carry arbitrary pickled
    objects need the explicit opt-in UC_UNSAFE_CHECKPOINTS=1, which restores the reference's `weights_only=False`."""
    import os

    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        if os.environ.get("UC_UNSAFE_CHECKPOINTS") == "1":
            return torch.load(path, map_location="cpu", weights_only=False)
        raise


def load_uniception_checkpoint(module: torch.nn.Module, path_or_ckpt, strict: bool = True):
    """Load a UniCeption-format checkpoint `{"model": state_dict, ...}` (croco.py:101-111, dust3r.py:206-209)."""
    ckpt = load_checkpoint_file(path_or_ckpt) if isinstance(path_or_ckpt, str) else path_or_ckpt
    return module.load_state_dict(_model(ckpt), strict=strict)
