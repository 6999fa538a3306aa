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
               "uc_head_post_fwd_params": _lib.HeadPostFwdParams, "uc_head_post_bwd_params": _lib.HeadPostBwdParams,
               "uc_headnorm_params": _lib.HeadNormParams, "uc_conv3x3_params": _lib.Conv3x3Params,
               "uc_patch_embed_params": _lib.PatchEmbedParams}
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


def test_argument_validation_returns_error_codes_without_touching_a_gpu():
    """The boundary's error behaviour (include/uc_b200.h:11-17): bad arguments are refused BEFORE any CUDA call with a
    negative UC_ERR_* code and a message in uc_last_error -- the counterpart of the native op's TORCH_CHECKs
    (curope.cpp:54-59).  Runs on a GPU-less host: nothing is launched."""
    import ctypes as C

    from uniception_b200 import _lib

    L = _lib.lib
    UC_ERR_BAD_SHAPE = -1
    before = _lib.launch_count()
    fake = 0x1000  # never dereferenced: validation fails first

    p = _lib.GemmParams()  # all-zero: null operands
    assert L.uc_gemm(C.byref(p), None) < 0 and "uc_gemm" in _lib.last_error()

    r = _lib.Rope2dParams(fake, fake, 1, 4, 2, 6, 48, 12, 6, 0, 100.0, 1.0)  # D = 6 is not a multiple of 4 (kernels.cu:91-94)
    assert L.uc_rope2d(C.byref(r), None) == UC_ERR_BAD_SHAPE and "multiple of 4" in _lib.last_error()
    r = _lib.Rope2dParams(None, fake, 1, 4, 2, 64, 512, 128, 64, 0, 100.0, 1.0)
    assert L.uc_rope2d(C.byref(r), None) == UC_ERR_BAD_SHAPE and "null" in _lib.last_error()

    h = _lib.HeadNormParams(fake, fake, 100, 128, fake, fake, None, None, None, None, 8, 2, 1e-6)  # ldx not a multiple of 8 / < heads*64
    assert L.uc_headnorm_fwd(C.byref(h), None) == UC_ERR_BAD_SHAPE and "head_dim is 64" in _lib.last_error()
    h = _lib.HeadNormParams(fake, fake, 128, 128, fake, fake, None, None, fake, None, 8, 2, 1e-6)  # positions without a table
    assert L.uc_headnorm_fwd(C.byref(h), None) == UC_ERR_BAD_SHAPE and "go together" in _lib.last_error()
    h = _lib.HeadNormParams(fake, fake, 128, 128, fake, None, None, None, None, None, 8, 2, 1e-6)  # bwd needs dgamma / dbeta
    assert L.uc_headnorm_bwd(C.byref(h), None) == UC_ERR_BAD_SHAPE

    assert L.uc_layerscale_fwd(fake, None, fake, fake, 4, 12, None) == UC_ERR_BAD_SHAPE  # cols % 8 != 0
    assert L.uc_layerscale_bwd(fake, fake, fake, fake, None, 4, 16, None) == UC_ERR_BAD_SHAPE  # dgamma missing
    assert L.uc_colsum(fake, 0, 12, 4, 12, fake, None) == UC_ERR_BAD_SHAPE and "uc_colsum" in _lib.last_error()
    assert L.uc_cast_bf16(None, fake, 8, None) == UC_ERR_BAD_SHAPE
    assert L.uc_rope2d_table(None, 4, 16, 100.0, 1.0, None) == UC_ERR_BAD_SHAPE
    assert _lib.launch_count() == before  # nothing reached a launch
    with pytest.raises(RuntimeError, match="libuc_b200 error -1"):
        _lib.check(UC_ERR_BAD_SHAPE)
