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
