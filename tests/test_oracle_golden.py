"""CPU: the oracle restatement reproduces every golden vector produced by the REAL reference
(`oracle/make_golden.py`), plus the C restatement of the native RoPE loop."""
import ctypes
import os

import numpy as np
import pytest
import torch

import dust3r_oracle as O
from golden_utils import load, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _close(a, b, rel=2e-5):
    ma, r = O.parity(a, b)
    assert r <= rel, (ma, r)


def test_index_ops_bit_exact():
    cfg, a = load("index_ops")
    assert torch.equal(O.patch_positions(2, 3, 5), a["positions_2_3_5"])
    assert torch.equal(O.pixel_shuffle(a["pixel_shuffle_in"], 2), a["pixel_shuffle_out"])
    for (n, ind), take in zip(cfg["take_cases"], cfg["take"]):
        assert O.feature_take_indices(n, ind)[0] == take
    i1, i2 = O.interleave(a["inter_a"], a["inter_b"])
    assert torch.equal(i1, a["inter_1"]) and torch.equal(i2, a["inter_2"])
    for (s1, s2), r in zip(cfg["sym_cases"], cfg["sym"]):
        assert O.is_symmetrized(s1, s2) == r


def test_rope2d_golden():
    cfg, a = load("rope2d")
    _close(O.rope2d(a["tokens"], a["positions"], cfg["base"], 1.0), a["out"], 1e-6)
    _close(O.rope2d(a["grad_out"], a["positions"], cfg["base"], -1.0), a["grad_in"], 1e-6)
    if "out_native_cpu" in a:  # the reference's own native CPU loop (oracle/_ref)
        _close(O.rope2d(a["tokens"], a["positions"], cfg["base"], 1.0), a["out_native_cpu"], 1e-6)
    # round trip: forward then F0=-1 restores the tokens (curope2d.py:24-28)
    rt = O.rope2d(O.rope2d(a["tokens"], a["positions"], 100.0, 1.0), a["positions"], 100.0, -1.0)
    _close(rt, a["tokens"], 1e-6)


def test_rope2d_c_restatement():
    lib_path = os.path.join(ROOT, "oracle", "librope2d_oracle.so")
    if not os.path.exists(lib_path):
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "librope2d_oracle.so"])
    lib = ctypes.CDLL(lib_path)
    cfg, a = load("rope2d")
    tok = a["tokens"].transpose(1, 2).contiguous().clone()  # [B,N,H,D]
    pos = a["positions"].contiguous()
    B, N, H, D = tok.shape
    lib.uc_oracle_rope2d(ctypes.c_void_p(tok.data_ptr()), ctypes.c_void_p(pos.data_ptr()), B, N, H, D,
                         ctypes.c_float(100.0), ctypes.c_float(1.0))
    _close(tok.transpose(1, 2), a["out"], 1e-6)
    p2 = torch.empty(2, 15, 2, dtype=torch.int64)
    lib.uc_oracle_positions(ctypes.c_void_p(p2.data_ptr()), 2, 3, 5)
    _, idx = load("index_ops")
    assert torch.equal(p2, idx["positions_2_3_5"])
    x = idx["pixel_shuffle_in"].contiguous()
    out = torch.empty_like(idx["pixel_shuffle_out"])
    lib.uc_oracle_pixel_shuffle(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), 2, 4, 3, 2, 2)
    assert torch.equal(out, idx["pixel_shuffle_out"])


@pytest.mark.parametrize("name", ["encoder_tiny", "encoder_tiny_ifr", "encoder_vitb16_224"])
def test_encoder_golden(name):
    cfg, a = load(name)
    sd = weights(cfg, "encoder.")
    if cfg["indices"] is None:
        _close(O.croco_encoder(sd, "encoder.", a["img"], cfg["depth"], cfg["heads"]), a["features"])
    else:
        f, inter = O.croco_encoder(sd, "encoder.", a["img"], cfg["depth"], cfg["heads"], indices=cfg["indices"])
        _close(f, a["features"])
        for i, t in enumerate(inter):
            _close(t, a[f"inter{i}"])


@pytest.mark.parametrize("name", ["dust3r_tiny_linear", "dust3r_tiny_linear_sym"])
def test_dust3r_linear_golden(name):
    cfg, a = load(name)
    sd = {k: v.requires_grad_(True) for k, v in weights(cfg).items()}
    r1, r2 = O.dust3r_forward(sd, a["img1"], a["img2"], enc_depth=cfg["enc_depth"], enc_heads=cfg["enc_heads"],
                              dec_depth=cfg["dec_depth"], dec_heads=cfg["dec_heads"], head="linear",
                              instances=(cfg["inst1"], cfg["inst2"]))
    _close(r1["pts3d"], a["pts3d_1"])
    _close(r1["conf"], a["conf_1"])
    _close(r2["pts3d_in_other_view"], a["pts3d_2"])
    _close(r2["conf"], a["conf_2"])
    O.bench_loss(r1, r2).backward()
    _close(sd["encoder.enc_blocks.0.attn.qkv.weight"].grad, a["grad_qkv0"], 1e-4)
    _close(sd["encoder.patch_embed.proj.weight"].grad, a["grad_patch"], 1e-4)
    _close(sd["info_sharing.multi_view_branches.1.0.cross_attn.projk.weight"].grad, a["grad_projk"], 1e-4)
    dig = np.array([[float(sd[k].grad.double().sum()), float(sd[k].grad.double().norm())] for k in cfg["grad_keys"]])
    ref = a["grad_digest"].numpy()
    assert np.allclose(dig[:, 1], ref[:, 1], rtol=2e-4, atol=1e-6)
