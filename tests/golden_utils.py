"""Helpers shared by CPU and GPU tests: load a golden fixture and re-derive its seeded weights."""
import json
import os

import numpy as np
import torch

import dust3r_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "cfg"}
    return cfg, arrays


def weights(cfg, prefix=""):
    sd = O.seeded_state_dict({k: tuple(v) for k, v in cfg["shapes"].items()}, cfg["seed"])
    return {prefix + k: v for k, v in sd.items()}


def c5_modules(cfg):
    """The configs[4]-style pipeline in miniature (patch-14 ViT IFR encoder -> DPT -> regressor) built from OUR modules,
    as one container whose state-dict keys (incl. the DPT aliases) must equal the reference's."""
    import torch.nn as nn

    import uniception_b200 as U
    from uniception_b200.prediction_heads import DPTFeature, DPTRegressionProcessor

    m = nn.Module()
    m.encoder = U.CroCoIntermediateFeatureReturner(
        name="enc", data_norm_type="dust3r", img_size=tuple(cfg["hw"]), patch_size=cfg["patch"], enc_embed_dim=cfg["C"],
        enc_depth=cfg["depth"], enc_num_heads=cfg["heads"], indices=list(cfg["indices"]), intermediates_only=True)
    m.dpt_feature_head = DPTFeature(patch_size=cfg["patch"], hooks=[0, 1, 2, 3], input_feature_dims=[cfg["C"]] * 4,
                                    layer_dims=[12, 24, 48, 96], feature_dim=32)
    m.dpt_regressor_head = DPTRegressionProcessor(input_feature_dim=32, output_dim=1)
    return m


def token_levels(cfg, a, dev="cpu"):
    """additional-token fixtures: [(maps, global extras or None, per-view extras or None)] for the final output and every
    tapped depth, plus the per-level loss weights used by oracle/make_golden.py:golden_additional_tokens."""
    n_levels = 1 + (len(cfg["indices"]) if cfg.get("indices") else 0)
    V = cfg["V"]
    levels = []
    for k in range(n_levels):
        maps = [a[f"l{k}_out{v}"].to(dev) for v in range(V)]
        ex = a[f"l{k}_extra"].to(dev) if cfg["T"] else None
        pv = [a[f"l{k}_pv{v}"].to(dev) for v in range(V)] if cfg["Tv"] else None
        levels.append((maps, ex, pv))
    return levels, [1.0 if k == 0 else k - 0.5 for k in range(n_levels)]


def token_loss(levels, weights_):
    tot = 0
    for (maps, ex, pv), wk in zip(levels, weights_):
        lv = sum(t.sum() for t in maps)
        if ex is not None:
            lv = lv + 2 * ex.sum()
        if pv is not None:
            lv = lv + 3 * sum(t.sum() for t in pv)
        tot = tot + wk * lv
    return tot
