"""GPU (-m gpu): per-kernel parity of the C-ABI entry points against fp32 oracle math on the SAME
bf16 inputs (tolerances written next to each check in tools/probe.py: rel-L2 <= 5e-3 for bf16
outputs, <= 5e-5 for fp32 outputs, bit-exact for index/gather ops)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["gemm_tn", "gemm_dgrad", "gemm_wgrad", "gemm_epilogues", "elementwise", "attn_fwd", "attn_bwd", "dpt_ops", "block_options", "conv3x3", "patch_embed"])
def test_kernel_parity(name):
    import probe

    assert getattr(probe, "check_" + name)(), f"{name}: parity outside tolerance (see stdout)"


def test_native_library_is_loaded_and_counts_launches():
    import torch

    from uniception_b200 import _lib, ops

    before = _lib.launch_count()
    ops.cast_bf16(torch.ones(1024, device="cuda"))
    assert _lib.launch_count() == before + 1
    assert any("libuc_b200.so" in l for l in open("/proc/self/maps"))


def test_rope2d_matches_reference_native_api_and_golden():
    """uc_rope2d through the reference's plugin interface (curope2d.py:31-39) on the committed golden."""
    import torch

    import dust3r_oracle as O
    from golden_utils import load
    from uniception_b200 import RoPE2D

    cfg, a = load("rope2d")
    tok = a["tokens"].cuda().clone().requires_grad_(True)
    pos = a["positions"].cuda()
    rope = RoPE2D(freq=cfg["base"])
    t_in = tok * 1.0  # non-leaf: the op is in place (mark_dirty), like cuRoPE2D
    out = rope(t_in, pos)
    ma, rel = O.parity(out, a["out"])
    assert rel <= 1e-6, (ma, rel)
    out.backward(a["grad_out"].cuda())
    ma, rel = O.parity(tok.grad, a["grad_in"])
    assert rel <= 1e-6, (ma, rel)
    with pytest.raises(RuntimeError):  # TORCH_CHECK-style argument errors (curope.cpp:54-59)
        rope(torch.zeros(1, 2, 3, 64, device="cuda"), torch.zeros(2, 3, 2, device="cuda", dtype=torch.int64))
