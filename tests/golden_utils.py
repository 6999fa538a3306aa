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
