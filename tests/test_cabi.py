"""CPU: the C-ABI library loads on a GPU-less host and exports every symbol include/uc_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "uc_b200.h")).read()
    return sorted(set(re.findall(r"UC_API\s+[\w\s\*]+?\b(uc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_api():
    names = _declared()
    assert "uc_gemm" in names and "uc_attn_fwd" in names and "uc_attn_bwd" in names and "uc_rope2d" in names
    assert len(names) >= 17


def test_library_exports_every_declared_symbol():
    from uniception_b200 import _lib

    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in include/uc_b200.h but not exported"
        assert name in _lib.EXPORTS, f"{name} has no ctypes prototype in uniception_b200/_lib.py"
    assert _lib.lib.uc_version() >= 100


def test_ctypes_struct_sizes_match_header_layout():
    """Field order / padding of the ctypes mirrors: compile a tiny C program against the header."""
    import subprocess
    import tempfile

    from uniception_b200 import _lib

    structs = {"uc_gemm_params": _lib.GemmParams, "uc_rope2d_params": _lib.Rope2dParams,
               "uc_layernorm_fwd_params": _lib.LayerNormFwdParams, "uc_layernorm_bwd_params": _lib.LayerNormBwdParams,
               "uc_attn_fwd_params": _lib.AttnFwdParams, "uc_attn_bwd_params": _lib.AttnBwdParams,
               "uc_head_post_fwd_params": _lib.HeadPostFwdParams, "uc_head_post_bwd_params": _lib.HeadPostBwdParams}
    prog = '#include <stdio.h>\n#include "uc_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = dict(zip(out[::2], map(int, out[1::2])))
    for n, cls in structs.items():
        assert ctypes.sizeof(cls) == sizes[n], (n, ctypes.sizeof(cls), sizes[n])


def test_compute_call_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from uniception_b200 import ops

    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(64, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16),
                 torch.zeros(64, 64, dtype=torch.bfloat16))
