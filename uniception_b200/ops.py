"""Raw (non-autograd) operator layer: torch tensors in, one C-ABI call each.

PyTorch is used for device memory and streams only; every function below launches hand-written
sm_100a kernels from `libuc_b200.so` on `torch.cuda.current_stream()`.  No function here has a
CPU or library fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib as L

PROFILE = None  # bench.py: set to a list to record (start_event, end_event, flops) per uc_gemm launch

_DT = {torch.bfloat16: L.UC_DTYPE_BF16, torch.float32: L.UC_DTYPE_F32, torch.float16: L.UC_DTYPE_F16}


_dev = None  # device of the tensors of the call in flight (set by _cuda)


def _stream():
    """The current stream OF THE TENSORS' DEVICE (a model on cuda:1 while cuda:0 is current must not launch on cuda:0)."""
    return C.c_void_p(torch.cuda.current_stream(_dev).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _cuda(*ts):
    global _dev
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("uniception_b200: tensors must live on a CUDA device (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"uniception_b200: tensors on different devices ({dev} vs {t.device})")
    if dev is not None:
        if dev.index != torch.cuda.current_device():
            raise RuntimeError(f"uniception_b200: tensors live on {dev} but the current CUDA device is {torch.cuda.current_device()}; "
                               "wrap the call in torch.cuda.device(...) (kernels launch on the current device)")
        _dev = dev


def gemm(a, b, out, *, a_layout=0, b_layout=0, bias=None, residual=None, aux_out=None, aux_in=None,
         positions=None, rope_table=None, rope_cols=0, gelu=False, gelu_bwd=False, atomic=False, split_k=0,
         relu=False, relu_bwd=False, c_colsum=None):
    """out[m,n] = epilogue(sum_k A(m,k) B(n,k)); see include/uc_b200.h (uc_gemm).
    c_colsum (fp32 [n], optional) += column sums of the bf16 `out` written by this call.
    a: [m,k] (a_layout 0) or [k,m] (1); b: [n,k] (0) or [k,n] (1); bf16, inner dim contiguous."""
    _cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, "uc_gemm operands are bf16"
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    m, k = (a.shape if a_layout == 0 else (a.shape[1], a.shape[0]))
    n, kb = (b.shape if b_layout == 0 else (b.shape[1], b.shape[0]))
    assert k == kb, f"contraction mismatch {k} vs {kb}"
    assert out.shape[0] == m and out.shape[1] == n, f"out {tuple(out.shape)} != ({m},{n})"
    epi = 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n
        epi |= L.EPI_BIAS
    if rope_table is not None:
        assert positions is not None and positions.dtype == torch.int32
        epi |= L.EPI_ROPE
    if gelu:
        assert aux_out is not None and aux_out.stride(0) == out.stride(0)
        epi |= L.EPI_GELU
    if gelu_bwd:
        assert aux_in is not None and aux_in.stride(0) == out.stride(0)
        epi |= L.EPI_GELU_BWD
    if relu:
        epi |= L.EPI_RELU
    if relu_bwd:
        assert aux_in is not None and aux_in.stride(0) == out.stride(0)
        epi |= L.EPI_RELU_BWD
    if residual is not None:
        assert residual.stride(0) == out.stride(0) and residual.stride(1) == 1
        epi |= L.EPI_RESIDUAL
        if residual.dtype == torch.float32:  # fp32 residual stream: fp32 in, fp32 out
            assert out.dtype == torch.float32, "an fp32 residual needs an fp32 output"
            epi |= L.EPI_RESIDUAL_F32
        else:
            assert residual.dtype == torch.bfloat16
    if atomic:
        epi |= L.EPI_ATOMIC
    p = L.GemmParams(
        _ptr(a), _ptr(b), _ptr(out), m, n, k, a_layout, b_layout, a.stride(0), b.stride(0), out.stride(0),
        _DT[out.dtype], epi, split_k, rope_cols,
        _ptr(bias), _ptr(residual), _ptr(aux_out), _ptr(aux_in), _ptr(positions), _ptr(rope_table), _ptr(c_colsum),
    )
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.uc_gemm(C.byref(p), _stream()))
        e1.record()
        PROFILE.append((e0, e1, 2.0 * m * n * k, (m, n, k, a_layout, b_layout, epi)))
        return out
    L.check(L.lib.uc_gemm(C.byref(p), _stream()))
    return out


def rope2d_(tokens_bnhd: torch.Tensor, positions: torch.Tensor, base: float, fwd: float):
    """In place on a [B,N,H,D] tensor (any strides with D contiguous) == curope.rope_2d (curope.cpp:49-69)."""
    _cuda(tokens_bnhd, positions)
    # same checks as the reference's TORCH_CHECKs (curope.cpp:54-59)
    if tokens_bnhd.dim() != 4:
        raise RuntimeError("tokens must have 4 dimensions")
    if positions.dim() != 3:
        raise RuntimeError("positions must have 3 dimensions")
    B, N, H, D = tokens_bnhd.shape
    if positions.shape[0] != B or positions.shape[1] != N:
        raise RuntimeError("batch size / seq_length differs between tokens & positions")
    if positions.shape[2] != 2:
        raise RuntimeError("positions.shape[2] must be equal to 2")
    if tokens_bnhd.stride(3) != 1:
        raise RuntimeError("tokens.stride(3) must be equal to 1")
    if positions.dtype != torch.int64:
        raise RuntimeError("positions must be int64")
    positions = positions.contiguous()
    p = L.Rope2dParams(_ptr(tokens_bnhd), _ptr(positions), B, N, H, D, tokens_bnhd.stride(0), tokens_bnhd.stride(1),
                       tokens_bnhd.stride(2), _DT[tokens_bnhd.dtype], float(base), float(fwd))
    L.check(L.lib.uc_rope2d(C.byref(p), _stream()))
    return tokens_bnhd


def rope2d_table(num_pos: int, base: float, f0: float, device, q: int = 16) -> torch.Tensor:
    t = torch.empty(num_pos, q, 2, dtype=torch.float32, device=device)
    _cuda(t)
    L.check(L.lib.uc_rope2d_table(_ptr(t), num_pos, q, float(base), float(f0), _stream()))
    return t


def layernorm_fwd(x, gamma, beta, eps, out_dtype=torch.bfloat16, save_stats=True):
    _cuda(x, gamma, beta)
    C_ = x.shape[-1]
    x2 = x.reshape(-1, C_)
    assert x2.is_contiguous()
    rows = x2.shape[0]
    y = torch.empty_like(x2, dtype=out_dtype)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
    p = L.LayerNormFwdParams(_ptr(x2), _ptr(y), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), rows, C_,
                             _DT[x2.dtype], _DT[out_dtype], float(eps))
    L.check(L.lib.uc_layernorm_fwd(C.byref(p), _stream()))
    return y.view(x.shape), mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None, dx_colsum=None):
    """dx (bf16) = LN'(dy) [+ dres]; dgamma/dbeta (fp32 [C]) are accumulated in place; dx_colsum (fp32 [C], optional)
    += column sums of dx (the bias gradient of the Linear that consumes dx as its output gradient)."""
    _cuda(dy, x, gamma)
    C_ = x.shape[-1]
    x2, dy2 = x.reshape(-1, C_), dy.reshape(-1, C_)
    assert x2.is_contiguous() and dy2.is_contiguous()
    if dres is not None:
        assert dres.dtype == torch.bfloat16 and dres.is_contiguous()
    dx = torch.empty_like(x2, dtype=torch.bfloat16)
    p = L.LayerNormBwdParams(_ptr(dy2), _ptr(x2), _ptr(dres), _ptr(dx), _ptr(gamma), _ptr(mean), _ptr(rstd),
                             _ptr(dgamma), _ptr(dbeta), x2.shape[0], C_, _DT[dy2.dtype], _DT[x2.dtype], _ptr(dx_colsum))
    L.check(L.lib.uc_layernorm_bwd(C.byref(p), _stream()))
    return dx.view(x.shape)


def headnorm_fwd(x, y, gamma, beta, eps, positions=None, rope_table=None):
    """qk_norm (+ fused RoPE): x, y bf16 column slices [rows, H*64] of packed projection buffers; gamma/beta fp32 [64]."""
    _cuda(x, y, gamma, beta)
    assert x.dtype == y.dtype == torch.bfloat16 and x.stride(1) == 1 and y.stride(1) == 1 and x.shape == y.shape
    assert x.shape[1] % 64 == 0 and gamma.numel() == 64 and beta.numel() == 64
    p = L.HeadNormParams(_ptr(x), _ptr(y), x.stride(0), y.stride(0), _ptr(gamma), _ptr(beta), None, None,
                         _ptr(positions), _ptr(rope_table), x.shape[0], x.shape[1] // 64, float(eps))
    L.check(L.lib.uc_headnorm_fwd(C.byref(p), _stream()))
    return y


def headnorm_bwd(g, x, gamma, dgamma, dbeta, eps):
    """In place on g: gradient w.r.t. the normalised (un-rotated) q or k -> gradient w.r.t. the raw projection x;
    dgamma / dbeta (fp32 [64]) are accumulated."""
    _cuda(g, x, gamma, dgamma, dbeta)
    assert x.dtype == g.dtype == torch.bfloat16 and x.stride(1) == 1 and g.stride(1) == 1 and x.shape == g.shape
    assert x.shape[1] % 64 == 0 and gamma.numel() == 64 and dgamma.dtype == dbeta.dtype == torch.float32
    p = L.HeadNormParams(_ptr(x), _ptr(g), x.stride(0), g.stride(0), _ptr(gamma), None, _ptr(dgamma), _ptr(dbeta),
                         None, None, x.shape[0], x.shape[1] // 64, float(eps))
    L.check(L.lib.uc_headnorm_bwd(C.byref(p), _stream()))
    return g


def layerscale_fwd(z, res, gamma):
    """res + gamma * z (bf16 [rows, C] contiguous; gamma fp32 [C]); res may be None."""
    _cuda(z, gamma)
    assert z.dtype == torch.bfloat16 and z.is_contiguous() and gamma.dtype == torch.float32 and gamma.numel() == z.shape[-1]
    assert res is None or (res.dtype == torch.bfloat16 and res.is_contiguous() and res.shape == z.shape)
    out = torch.empty_like(z)
    z2 = z.reshape(-1, z.shape[-1])
    L.check(L.lib.uc_layerscale_fwd(_ptr(z), _ptr(res), _ptr(gamma), _ptr(out), z2.shape[0], z2.shape[1], _stream()))
    return out


def layerscale_bwd(dy, z, gamma, dgamma):
    """dz = gamma * dy (returned, bf16); dgamma (fp32 [C]) += column sums of dy * z."""
    _cuda(dy, z, gamma, dgamma)
    assert dy.dtype == z.dtype == torch.bfloat16 and dy.is_contiguous() and z.is_contiguous() and dy.shape == z.shape
    dz = torch.empty_like(dy)
    z2 = z.reshape(-1, z.shape[-1])
    L.check(L.lib.uc_layerscale_bwd(_ptr(dy), _ptr(z), _ptr(gamma), _ptr(dz), _ptr(dgamma), z2.shape[0], z2.shape[1], _stream()))
    return dz


# UC_ATTN_BWD=1 (default): fused one-kernel backward with an fp32 dQ accumulator; UC_ATTN_BWD=3: two bit-reproducible kernels
_ATTN_BWD_V1 = __import__("os").environ.get("UC_ATTN_BWD", "1") != "3"


def attn_fwd(q, k, v, B, H, Nq, Nk, scale, out=None):
    """q: [B*Nq, >=H*64] bf16 view (row stride = ld), k/v: [B*Nk, ...]; returns (o [B*Nq, H*64], lse [B,H,Nq])."""
    _cuda(q, k, v)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    assert q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1
    o = torch.empty(B * Nq, H * 64, dtype=torch.bfloat16, device=q.device) if out is None else out
    lse = torch.empty(B, H, Nq, dtype=torch.float32, device=q.device)
    p = L.AttnFwdParams(_ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(lse), B, H, Nq, Nk,
                        q.stride(0), k.stride(0), v.stride(0), o.stride(0), float(scale))
    L.check(L.lib.uc_attn_fwd(C.byref(p), _stream()))
    return o, lse


def attn_bwd(q, k, v, o, d_o, lse, B, H, Nq, Nk, scale, dq, dk, dv, q_positions=None, k_positions=None, rope_table=None):
    """Writes dq/dk/dv (bf16 views with the same addressing as q/k/v).  With positions + table the
    gradients are returned w.r.t. the *un-rotated* projections (inverse RoPE fused)."""
    _cuda(q, k, v, o, d_o, lse, dq, dk, dv)
    assert d_o.dtype == torch.bfloat16 and d_o.stride(0) == o.stride(0) and d_o.stride(1) == 1
    # workspace: per (b, h) the per-query statistics padded to 64-query tiles, [tile][lse*log2e | delta][64]
    delta = torch.empty(B * H * ((Nq + 63) // 64) * 128, dtype=torch.float32, device=q.device)
    dq_acc = torch.empty(B * Nq, H * 64, dtype=torch.float32, device=q.device) if _ATTN_BWD_V1 else None
    p = L.AttnBwdParams(_ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(d_o), _ptr(lse), _ptr(delta), _ptr(dq_acc),
                        _ptr(dq), _ptr(dk), _ptr(dv), B, H, Nq, Nk,
                        q.stride(0), k.stride(0), v.stride(0), o.stride(0), dq.stride(0), dk.stride(0), dv.stride(0),
                        float(scale), _ptr(q_positions), _ptr(k_positions), _ptr(rope_table))
    L.check(L.lib.uc_attn_bwd(C.byref(p), _stream()))
    return dq, dk, dv


def patchify(img: torch.Tensor, patch: int) -> torch.Tensor:
    _cuda(img)
    assert img.dtype == torch.float32 and img.is_contiguous()
    B, Cc, H, W = img.shape
    assert H % patch == 0, f"Input image height ({H}) is not a multiple of patch size ({patch})."
    assert W % patch == 0, f"Input image width ({W}) is not a multiple of patch size ({patch})."
    k = Cc * patch * patch
    kp = k if patch % 8 == 0 else (k + 63) // 64 * 64  # odd patch sizes: row pitch padded to the GEMM granularity (zeros)
    cols = torch.empty(B * (H // patch) * (W // patch), kp, dtype=torch.bfloat16, device=img.device)
    L.check(L.lib.uc_patchify(_ptr(img), _ptr(cols), B, Cc, H, W, patch, _stream()))
    return cols


def colsum_(x: torch.Tensor, out: torch.Tensor):
    """out[cols] (fp32) += column sums of x [rows, cols]."""
    _cuda(x, out)
    assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32
    L.check(L.lib.uc_colsum(_ptr(x), _DT[x.dtype], x.stride(0), x.shape[0], x.shape[1], _ptr(out), _stream()))
    return out


def cast_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    dst = torch.empty_like(src, dtype=torch.bfloat16) if out is None else out
    L.check(L.lib.uc_cast_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()))
    return dst


def nlc_to_nchw(x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[B, L, C] (bf16/fp32) -> fp32 [B, C, h, w] (encoders/croco.py:177-180)."""
    _cuda(x)
    B, Lh, Cc = x.shape
    assert x.is_contiguous() and Lh == h * w
    out = torch.empty(B, Cc, h, w, dtype=torch.float32, device=x.device)
    L.check(L.lib.uc_nlc_to_nchw(_ptr(x), _DT[x.dtype], _ptr(out), B, Lh, Cc, _stream()))
    return out


def nchw_to_nlc(x: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """fp32 [B, C, h, w] -> [B, h*w, C] (info_sharing/cross_attention_transformer.py:222-225)."""
    _cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    B, Cc, h, w = x.shape
    out = torch.empty(B, h * w, Cc, dtype=dtype, device=x.device)
    L.check(L.lib.uc_nchw_to_nlc(_ptr(x), _ptr(out), _DT[dtype], B, h * w, Cc, _stream()))
    return out


def head_post_fwd(y: torch.Tensor, B: int, h: int, w: int, patch: int, conf_min=1.0, conf_max=math.inf):
    _cuda(y)
    assert y.dtype == torch.float32 and y.is_contiguous() and y.shape[0] == B * h * w and y.shape[1] >= 4 * patch * patch
    pts = torch.empty(B, h * patch, w * patch, 3, dtype=torch.float32, device=y.device)
    conf = torch.empty(B, h * patch, w * patch, 1, dtype=torch.float32, device=y.device)
    p = L.HeadPostFwdParams(_ptr(y), _ptr(pts), _ptr(conf), B, h, w, patch, float(conf_min), float(conf_max), y.stride(0))
    L.check(L.lib.uc_head_post_fwd(C.byref(p), _stream()))
    return pts, conf


def head_post_bwd(y, dpts, dconf, B, h, w, patch, conf_min=1.0, conf_max=math.inf, dtype=torch.bfloat16):
    _cuda(y, dpts, dconf)
    dpts, dconf = dpts.contiguous(), dconf.contiguous()
    dense = y.shape[1] == 4 * patch * patch
    dy = torch.empty(y.shape, dtype=dtype, device=y.device) if dense else torch.zeros(y.shape, dtype=dtype, device=y.device)
    p = L.HeadPostBwdParams(_ptr(y), _ptr(dpts), _ptr(dconf), _ptr(dy), _DT[dtype], B, h, w, patch,
                            float(conf_min), float(conf_max), y.stride(0))
    L.check(L.lib.uc_head_post_bwd(C.byref(p), _stream()))
    return dy


# ---- DPT head support (NHWC bf16 feature maps == token-major [B*H*W, C]) ----
def conv_out_hw(H: int, W: int, stride: int):
    return (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1


def softmax_rows_fwd(s: torch.Tensor, valid: int, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """s fp32 [rows, ld] -> bf16 [rows, ld]: softmax(scale * s[:, :valid]) with zeros in the padding columns."""
    _cuda(s)
    assert s.dtype == torch.float32 and s.is_contiguous() and s.dim() == 2
    p = torch.empty(s.shape, dtype=torch.bfloat16, device=s.device) if out is None else out
    assert p.dtype == torch.bfloat16 and p.is_contiguous() and p.shape == s.shape
    L.check(L.lib.uc_softmax_rows_fwd(_ptr(s), _ptr(p), s.shape[0], valid, s.shape[1], float(scale), _stream()))
    return p


def softmax_rows_bwd(p: torch.Tensor, dp: torch.Tensor, valid: int, scale: float) -> torch.Tensor:
    """dS (bf16) = scale * P o (dP - rowsum(P o dP)); P bf16, dP fp32, both [rows, ld]."""
    _cuda(p, dp)
    assert p.dtype == torch.bfloat16 and dp.dtype == torch.float32 and p.is_contiguous() and dp.is_contiguous() and p.shape == dp.shape
    ds = torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)
    L.check(L.lib.uc_softmax_rows_bwd(_ptr(p), _ptr(dp), _ptr(ds), p.shape[0], valid, p.shape[1], float(scale), _stream()))
    return ds


def patch_embed_ok(img: torch.Tensor, patch: int, n: int) -> bool:
    """Shapes uc_patch_embed accepts (TMA box constraints, include/uc_b200.h); everything else takes patchify + gemm."""
    return patch in (16, 32) and img.shape[-1] % 4 == 0 and n % 128 == 0 and img.shape[1] == 3


def patch_embed(img: torch.Tensor, w32: torch.Tensor, bias: Optional[torch.Tensor], patch: int) -> torch.Tensor:
    """im2col-free patch embedding (uc_patch_embed): img fp32 [B,3,H,W], w32 fp32 [n, 3*patch*patch] -> bf16 [B*Hp*Wp, n]."""
    _cuda(img, w32)
    assert img.dtype == torch.float32 and img.is_contiguous() and w32.dtype == torch.float32 and w32.is_contiguous()
    B, _, H, W = img.shape
    n = w32.shape[0]
    assert w32.shape[1] == 3 * patch * patch
    out = torch.empty(B * (H // patch) * (W // patch), n, dtype=torch.bfloat16, device=img.device)
    p = L.PatchEmbedParams(_ptr(img), _ptr(w32), _ptr(bias), _ptr(out), B, H, W, patch, n)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.uc_patch_embed(C.byref(p), _stream()))
        e1.record()
        PROFILE.append((e0, e1, 2.0 * out.shape[0] * n * 3 * patch * patch, ("patch_embed", out.shape[0], n, 3 * patch * patch, 0, 0)))
        return out
    L.check(L.lib.uc_patch_embed(C.byref(p), _stream()))
    return out


def _conv_call(p, B, H, W, cin, cout, mode):
    """uc_conv3x3, timed with CUDA events when bench.py's per-launch GEMM profile is on (the convolution runs on the same
    tcgen05 kernel as uc_gemm; algorithmic FLOPs = 2 * pixels * 9 * cin * cout)."""
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.uc_conv3x3(C.byref(p), _stream()))
        e1.record()
        PROFILE.append((e0, e1, 2.0 * B * H * W * 9 * cin * cout, ("conv3x3", mode, B * H * W, cin, cout, 0)))
        return
    L.check(L.lib.uc_conv3x3(C.byref(p), _stream()))


def conv3x3_fwd(x: torch.Tensor, w16: torch.Tensor, B: int, H: int, W: int, bias=None, relu: bool = False, residual=None) -> torch.Tensor:
    """Implicit-GEMM 3x3 / stride 1 / pad 1 convolution (uc_conv3x3 mode 0).  x bf16 [B*H*W, cin], w16 bf16 [cout, 9*cin]
    (tap-major), optional fp32 bias, fused ReLU or `+ residual` (bf16 [B*H*W, cout]).  Returns bf16 [B*H*W, cout]."""
    _cuda(x, w16)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[0] == B * H * W and w16.is_contiguous()
    cin, cout = x.shape[1], w16.shape[0]
    assert w16.shape[1] == 9 * cin
    y = torch.empty(B * H * W, cout, dtype=torch.bfloat16, device=x.device)
    p = L.Conv3x3Params(0, B, H, W, cin, cout, _ptr(x), _ptr(w16), _ptr(y), None, None, None, _ptr(bias), _ptr(residual), None, int(relu))
    _conv_call(p, B, H, W, cin, cout, 0)
    return y


def conv3x3_dgrad(dy: torch.Tensor, w16: torch.Tensor, B: int, H: int, W: int, relu_out=None) -> torch.Tensor:
    """dx of the same convolution (uc_conv3x3 mode 1); relu_out: dx *= (relu_out > 0) (the conv's input was a ReLU output)."""
    _cuda(dy, w16)
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and dy.shape[0] == B * H * W
    cout = w16.shape[0]
    cin = w16.shape[1] // 9
    assert dy.shape[1] == cout
    dx = torch.empty(B * H * W, cin, dtype=torch.bfloat16, device=dy.device)
    p = L.Conv3x3Params(1, B, H, W, cin, cout, None, _ptr(w16), None, _ptr(dy), _ptr(dx), None, None, None, _ptr(relu_out), 0)
    _conv_call(p, B, H, W, cin, cout, 1)
    return dx


def conv3x3_wgrad_(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, B: int, H: int, W: int) -> None:
    """dw (fp32 [cout, 9*cin]) += the weight gradient of the same convolution (uc_conv3x3 mode 2)."""
    _cuda(x, dy, dw)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dw.dtype == torch.float32
    assert x.is_contiguous() and dy.is_contiguous() and dw.is_contiguous()
    cin, cout = x.shape[1], dy.shape[1]
    assert tuple(dw.shape) == (cout, 9 * cin)
    p = L.Conv3x3Params(2, B, H, W, cin, cout, _ptr(x), None, None, _ptr(dy), None, _ptr(dw), None, None, None, 0)
    _conv_call(p, B, H, W, cin, cout, 2)


def im2col3x3(x: torch.Tensor, B: int, H: int, W: int, stride: int = 1) -> torch.Tensor:
    """x [B*H*W, C] bf16 -> cols [B*Ho*Wo, 9*C] (tap-major, pad 1)."""
    _cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[0] == B * H * W
    C_ = x.shape[1]
    Ho, Wo = conv_out_hw(H, W, stride)
    cols = torch.empty(B * Ho * Wo, 9 * C_, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib.uc_im2col3x3(_ptr(x), _ptr(cols), B, H, W, C_, stride, _stream()))
    return cols


def col2im3x3(dcols: torch.Tensor, B: int, H: int, W: int, stride: int = 1) -> torch.Tensor:
    _cuda(dcols)
    assert dcols.dtype == torch.bfloat16 and dcols.is_contiguous()
    C_ = dcols.shape[1] // 9
    dx = torch.empty(B * H * W, C_, dtype=torch.bfloat16, device=dcols.device)
    L.check(L.lib.uc_col2im3x3(_ptr(dcols), _ptr(dx), B, H, W, C_, stride, _stream()))
    return dx


def depth_to_space(x: torch.Tensor, B: int, h: int, w: int, s: int) -> torch.Tensor:
    """[B*h*w, s*s*C] (columns (i,j,c)) -> NHWC [B*(h*s)*(w*s), C]."""
    _cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    C_ = x.shape[1] // (s * s)
    out = torch.empty(B * h * s * w * s, C_, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib.uc_depth_space(_ptr(x), _ptr(out), B, h, w, C_, s, 1, _stream()))
    return out


def space_to_depth(x: torch.Tensor, B: int, h: int, w: int, s: int) -> torch.Tensor:
    _cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    C_ = x.shape[1]
    out = torch.empty(B * h * w, s * s * C_, dtype=torch.bfloat16, device=x.device)
    L.check(L.lib.uc_depth_space(_ptr(x), _ptr(out), B, h, w, C_, s, 0, _stream()))
    return out


def bilinear_fwd(x: torch.Tensor, B: int, Hi: int, Wi: int, Ho: int, Wo: int) -> torch.Tensor:
    """x bf16 or fp32 NHWC [B*Hi*Wi, C] -> bf16 [B*Ho*Wo, C] (align_corners=True)."""
    _cuda(x)
    assert x.dtype in (torch.bfloat16, torch.float32) and x.is_contiguous()
    out = torch.empty(B * Ho * Wo, x.shape[1], dtype=torch.bfloat16, device=x.device)
    fn = L.lib.uc_bilinear_fwd if x.dtype == torch.bfloat16 else L.lib.uc_bilinear_fwd_f32in
    L.check(fn(_ptr(x), _ptr(out), B, Hi, Wi, Ho, Wo, x.shape[1], _stream()))
    return out


def bilinear_bwd(dout: torch.Tensor, B: int, Hi: int, Wi: int, Ho: int, Wo: int) -> torch.Tensor:
    _cuda(dout)
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous()
    din = torch.empty(B * Hi * Wi, dout.shape[1], dtype=torch.bfloat16, device=dout.device)
    L.check(L.lib.uc_bilinear_bwd(_ptr(dout), _ptr(din), B, Hi, Wi, Ho, Wo, dout.shape[1], _stream()))
    return din


def elementwise(op: int, a, b=None, c=None):
    """0: a+b  1: relu(a)  2: a*(b>0)  3: a+b+c   (bf16, same shapes)"""
    _cuda(a, b, c)
    assert a.dtype == torch.bfloat16 and a.is_contiguous() and a.numel() % 8 == 0
    out = torch.empty_like(a)
    L.check(L.lib.uc_elementwise(op, _ptr(a), _ptr(b), _ptr(c), _ptr(out), a.numel(), _stream()))
    return out
