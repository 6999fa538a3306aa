"""Per-op autograd Functions over the C-ABI kernels, for using the blocks on their own.

The whole-module engines (engine.py) do not go through these; they schedule forward/backward by
hand.  Conventions: activations are bf16 (fp32 inputs are cast on entry), parameters are fp32
masters with a version-keyed bf16 operand cache, parameter gradients are returned to autograd in fp32.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import engine as E
from . import ops
from .rope import fusable_rope

def w16(param: torch.Tensor) -> torch.Tensor:
    """bf16 operand copy of an fp32 parameter, refreshed when the parameter changes.  The copy is cached ON the parameter
    object (a WeakKeyDictionary keyed by tensors compares keys with `==` on hash collisions, which is elementwise)."""
    key = (param.data_ptr(), param._version, str(param.device))
    hit = getattr(param, "_uc_w16", None)
    if hit is None or hit[0] != key:
        hit = (key, ops.cast_bf16(param.detach().contiguous()).view(param.shape[0], -1))
        try:
            param._uc_w16 = hit
        except AttributeError:  # pragma: no cover  (plain tensors without a __dict__)
            pass
    return hit[1]


def _act2d(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.bfloat16:
        x2 = x2.to(torch.bfloat16)
    return x2.contiguous()


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        x2 = _act2d(x)
        w = w16(weight)
        out = torch.empty(x2.shape[0], w.shape[0], dtype=torch.bfloat16, device=x2.device)
        r2 = _act2d(residual) if residual is not None else None
        ops.gemm(x2, w, out, bias=bias.detach() if bias is not None else None, residual=r2)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.in_shape = x.shape
        ctx.w_shape = weight.shape
        return out.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = _act2d(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(dy2.shape[0], w.shape[1], dtype=torch.bfloat16, device=dy2.device)
            ops.gemm(dy2, w, dx, b_layout=1)
            dx = dx.view(ctx.in_shape)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(w.shape, dtype=torch.float32, device=dy2.device)
            ops.gemm(dy2, x2, dw, a_layout=1, b_layout=1, atomic=True)
            dw = dw.view(ctx.w_shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(w.shape[0], dtype=torch.float32, device=dy2.device)
            ops.colsum_(dy2, db)
        return dx, dw, db, (dy if ctx.has_res else None)


class MlpFn(torch.autograd.Function):
    """fc2(gelu(fc1 x)) [+ residual] with the GELU and its derivative fused into the GEMM epilogues."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        x2 = _act2d(x)
        w1b, w2b = w16(w1), w16(w2)
        act = torch.empty(x2.shape[0], w1b.shape[0], dtype=torch.bfloat16, device=x2.device)
        pre = torch.empty_like(act)
        ops.gemm(x2, w1b, act, bias=b1.detach(), gelu=True, aux_out=pre)
        out = torch.empty(x2.shape[0], w2b.shape[0], dtype=torch.bfloat16, device=x2.device)
        ops.gemm(act, w2b, out, bias=b2.detach(), residual=_act2d(residual) if residual is not None else None)
        ctx.save_for_backward(x2, w1b, w2b, pre, act)
        ctx.has_res = residual is not None
        ctx.in_shape = x.shape
        ctx.shapes = (w1.shape, w2.shape)
        return out.view(*x.shape[:-1], w2b.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w1b, w2b, pre, act = ctx.saved_tensors
        dy2 = _act2d(dy)
        dev = dy2.device
        d_pre = torch.empty_like(pre)
        ops.gemm(dy2, w2b, d_pre, b_layout=1, gelu_bwd=True, aux_in=pre)
        dw2 = torch.zeros(w2b.shape, dtype=torch.float32, device=dev)
        ops.gemm(dy2, act, dw2, a_layout=1, b_layout=1, atomic=True)
        db2 = torch.zeros(w2b.shape[0], dtype=torch.float32, device=dev)
        ops.colsum_(dy2, db2)
        dw1 = torch.zeros(w1b.shape, dtype=torch.float32, device=dev)
        ops.gemm(d_pre, x2, dw1, a_layout=1, b_layout=1, atomic=True)
        db1 = torch.zeros(w1b.shape[0], dtype=torch.float32, device=dev)
        ops.colsum_(d_pre, db1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x2.shape, dtype=torch.bfloat16, device=dev)
            ops.gemm(d_pre, w1b, dx, b_layout=1)
            dx = dx.view(ctx.in_shape)
        return dx, dw1.view(ctx.shapes[0]), db1, dw2.view(ctx.shapes[1]), db2, (dy if ctx.has_res else None)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        xin = x.contiguous()
        if xin.dtype not in (torch.bfloat16, torch.float32):
            xin = xin.to(torch.bfloat16)
        y, mean, rstd = ops.layernorm_fwd(xin, weight.detach(), bias.detach(), eps, torch.bfloat16)
        ctx.save_for_backward(xin, weight.detach(), mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        xin, weight, mean, rstd = ctx.saved_tensors
        dyc = dy.contiguous()
        if dyc.dtype not in (torch.bfloat16, torch.float32):
            dyc = dyc.to(torch.bfloat16)
        dg = torch.zeros_like(weight)
        db = torch.zeros_like(weight)
        dx = ops.layernorm_bwd(dyc, xin, weight, mean, rstd, dg, db)
        return dx.to(xin.dtype) if xin.dtype != torch.bfloat16 else dx, dg, db, None


class LayerScaleFn(torch.autograd.Function):
    """x * gamma (utils/transformer_blocks.py:389-412) on bf16 token tensors."""

    @staticmethod
    def forward(ctx, x, gamma):
        xin = x.contiguous() if x.dtype == torch.bfloat16 else x.to(torch.bfloat16).contiguous()
        ctx.save_for_backward(xin, gamma.detach())
        ctx.in_dtype = x.dtype
        return ops.layerscale_fwd(xin, None, gamma.detach())

    @staticmethod
    def backward(ctx, dy):
        xin, gamma = ctx.saved_tensors
        dyc = dy.contiguous() if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16).contiguous()
        dg = torch.zeros_like(gamma)
        dx = ops.layerscale_bwd(dyc, xin, gamma, dg)
        return dx if ctx.in_dtype == torch.bfloat16 else dx.to(ctx.in_dtype), dg


class HeadNormFn(torch.autograd.Function):
    """qk_norm on columns [off, off+C) of a packed projection: LayerNorm over every 64-wide head segment
    (utils/transformer_blocks.py:199-200, :222).  RoPE stays in `AttentionFn`, whose backward returns the gradient w.r.t. the
    un-rotated values this function produced."""

    @staticmethod
    def forward(ctx, src, off, C, weight, bias, eps):
        s2 = _act2d(src)
        out = torch.empty(s2.shape[0], C, dtype=torch.bfloat16, device=s2.device)
        ops.headnorm_fwd(s2[:, off:off + C], out, weight.detach().contiguous(), bias.detach().contiguous(), eps)
        ctx.save_for_backward(s2, weight.detach().contiguous())
        ctx.meta = (off, C, eps, src.shape, src.dtype)
        return out.view(*src.shape[:-1], C)

    @staticmethod
    def backward(ctx, dy):
        s2, w = ctx.saved_tensors
        off, C, eps, shp, dt = ctx.meta
        dsrc = torch.zeros(s2.shape, dtype=torch.bfloat16, device=s2.device)
        dsrc[:, off:off + C] = _act2d(dy)
        dg, db = torch.zeros_like(w), torch.zeros_like(w)
        ops.headnorm_bwd(dsrc[:, off:off + C], s2[:, off:off + C], w, dg, db, eps)
        return (dsrc.view(shp) if dt == torch.bfloat16 else dsrc.view(shp).to(dt)), None, None, dg, db, None


class AttentionFn(torch.autograd.Function):
    """softmax(q k^T / 8) v on token-major bf16 sources; optional fused 2-D RoPE (rotated copies in
    forward, inverse rotation fused into the backward kernels)."""

    @staticmethod
    def forward(ctx, q_src, kv_src, B, Nq, Nk, H, q_off, k_off, v_off, qpos32, kpos32, table, scale=0.125):
        C = H * 64
        q2, kv2 = _act2d(q_src), _act2d(kv_src)
        q = q2[:, q_off:q_off + C]
        k = kv2[:, k_off:k_off + C]
        v = kv2[:, v_off:v_off + C]
        if table is not None:
            q, k = q.contiguous().clone(), k.contiguous().clone()
            # rotate in place on [B,N,H,64] views with the int32->table free standalone kernel
            base, f0, _ = table
            ops.rope2d_(q.view(B, Nq, H, 64), qpos32.view(B, Nq, 2).long(), base, f0)
            ops.rope2d_(k.view(B, Nk, H, 64), kpos32.view(B, Nk, 2).long(), base, f0)
        o, lse = ops.attn_fwd(q, k, v, B, H, Nq, Nk, scale)
        ctx.scale = scale
        ctx.save_for_backward(q, k, v, o, lse, qpos32, kpos32, table[2] if table is not None else None)
        ctx.dims = (B, Nq, Nk, H, q_off, k_off, v_off, q2.shape, kv2.shape, q_src.shape, kv_src.shape, q_src is kv_src)
        return o.view(*q_src.shape[:-1], C)

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse, qpos32, kpos32, tab = ctx.saved_tensors
        B, Nq, Nk, H, q_off, k_off, v_off, q2s, kv2s, qs, kvs, same = ctx.dims
        C = H * 64
        d2 = _act2d(d_o)
        dq_src = torch.zeros(q2s, dtype=torch.bfloat16, device=d2.device)
        dkv_src = dq_src if same else torch.zeros(kv2s, dtype=torch.bfloat16, device=d2.device)
        ops.attn_bwd(q, k, v, o, d2, lse, B, H, Nq, Nk, ctx.scale, dq_src[:, q_off:q_off + C], dkv_src[:, k_off:k_off + C],
                     dkv_src[:, v_off:v_off + C], q_positions=qpos32 if tab is not None else None,
                     k_positions=kpos32 if tab is not None else None, rope_table=tab)
        return (dq_src.view(qs), None if same else dkv_src.view(kvs)) + (None,) * 11


class GeneralAttentionFn(torch.autograd.Function):
    """softmax(q k^T * scale) v for head dims the fused kernels do not cover (any multiples of 64; q/k width != v width
    allowed): the attention of the DiffAttention family (utils/transformer_blocks.py:686-945) -- 128-wide self-attention heads and
    64-wide q/k against 128-wide v at the default dim 768 / 12 heads.  UN-FUSED: per (batch, head) the scores are a uc_gemm
    (fp32 out), the softmax a row kernel (uc_softmax_rows_*), P V / the four gradient products uc_gemms again; P (bf16) is
    kept for the backward.  Keys are zero-padded to a multiple of 64 (GEMM n / k constraint), masked in the softmax.
    q4 [B,H,Nq,dqk], k4 [B,H,Nk,dqk], v4 [B,H,Nk,dv] -> [B,H,Nq,dv] bf16."""

    @staticmethod
    def forward(ctx, q4, k4, v4, scale):
        B, H, Nq, dqk = q4.shape
        Nk, dv = k4.shape[2], v4.shape[3]
        assert dqk % 64 == 0 and dv % 64 == 0, "head dims must be multiples of 64"
        if not q4.is_cuda:
            raise RuntimeError("uniception_b200: tensors must live on a CUDA device (no CPU fallback)")
        Nkp = (Nk + 63) // 64 * 64
        dev = q4.device
        q = q4.to(torch.bfloat16).contiguous()
        k = torch.zeros(B, H, Nkp, dqk, dtype=torch.bfloat16, device=dev)
        v = torch.zeros(B, H, Nkp, dv, dtype=torch.bfloat16, device=dev)
        k[:, :, :Nk] = k4
        v[:, :, :Nk] = v4
        P = torch.empty(B, H, Nq, Nkp, dtype=torch.bfloat16, device=dev)
        o = torch.empty(B, H, Nq, dv, dtype=torch.bfloat16, device=dev)
        s = torch.empty(Nq, Nkp, dtype=torch.float32, device=dev)
        for b in range(B):
            for h in range(H):
                ops.gemm(q[b, h], k[b, h], s)
                ops.softmax_rows_fwd(s, Nk, scale, out=P[b, h])
                ops.gemm(P[b, h], v[b, h], o[b, h], b_layout=1)
        ctx.save_for_backward(q, k, v, P)
        ctx.meta = (Nk, float(scale), q4.dtype, k4.dtype, v4.dtype)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, P = ctx.saved_tensors
        Nk, scale, tq, tk, tv = ctx.meta
        B, H, Nq, dqk = q.shape
        Nkp, dv = k.shape[2], v.shape[3]
        dev = q.device
        do = d_o.to(torch.bfloat16).contiguous()
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        dvp = torch.empty_like(v)
        dp = torch.empty(Nq, Nkp, dtype=torch.float32, device=dev)
        for b in range(B):
            for h in range(H):
                ops.gemm(P[b, h], do[b, h], dvp[b, h], a_layout=1, b_layout=1)  # dV = P^T dO
                ops.gemm(do[b, h], v[b, h], dp)                                  # dP = dO V^T
                ds = ops.softmax_rows_bwd(P[b, h], dp, Nk, scale)
                ops.gemm(ds, k[b, h], dq[b, h], b_layout=1)                      # dQ = dS K
                ops.gemm(ds, q[b, h], dk[b, h], a_layout=1, b_layout=1)          # dK = dS^T Q
        return dq.to(tq), dk[:, :, :Nk].to(tk), dvp[:, :, :Nk].to(tv), None


def general_attention(q4, k4, v4, scale):
    return GeneralAttentionFn.apply(q4, k4, v4, scale)


# ------------------------------------------------------------------------------------------------
# functional front-ends used by blocks.py
# ------------------------------------------------------------------------------------------------
def linear(x, weight, bias, residual=None, gelu=False):
    assert not gelu, "use mlp() for the fused GELU path"
    return LinearFn.apply(x, weight, bias, residual)


def mlp(x, fc1: nn.Linear, fc2: nn.Linear, residual=None):
    return MlpFn.apply(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias, residual)


def layer_norm(x, norm: nn.LayerNorm):
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


def head_norm(src, off, C, norm: nn.LayerNorm):
    return HeadNormFn.apply(src, off, C, norm.weight, norm.bias, norm.eps)


def layer_scale(x, gamma):
    return LayerScaleFn.apply(x, gamma)


def attention(q_src, kv_src, B, Nq, Nk, H, q_off, k_off, v_off, qpos=None, kpos=None, rope=None, scale=0.125):
    """scale: softmax scale (head_dim^-0.5, times the scalable-softmax / entropy-scaling query multipliers)."""
    fr = fusable_rope(rope)
    if rope is not None and fr is None:
        # arbitrary positional-encoding plugin: call it on [B,H,N,d] tensors under autograd, as the
        # reference does (utils/transformer_blocks.py:224-229), then run the un-rotated kernel
        C = H * 64
        q4 = q_src[..., q_off:q_off + C].reshape(B, Nq, H, 64).transpose(1, 2)
        k4 = kv_src[..., k_off:k_off + C].reshape(B, Nk, H, 64).transpose(1, 2)
        q4, k4 = rope(q4, qpos), rope(k4, kpos)
        q_r = q4.transpose(1, 2).reshape(B, Nq, C)
        kv_r = torch.cat((k4.transpose(1, 2).reshape(B, Nk, C), kv_src[..., v_off:v_off + C].reshape(B, Nk, C)), dim=-1)
        return AttentionFn.apply(q_r, kv_r, B, Nq, Nk, H, 0, 0, C, None, None, None, scale)
    if fr is None:
        return AttentionFn.apply(q_src, kv_src, B, Nq, Nk, H, q_off, k_off, v_off, None, None, None, scale)
    assert qpos is not None and kpos is not None
    base, f0 = fr
    num_pos = int(max(int(qpos.max()), int(kpos.max()))) + 1  # host sync: granular API only
    table = E.rope_table(num_pos, base, f0, q_src.device)
    q32 = qpos.reshape(-1, 2).to(torch.int32).contiguous()
    k32 = kpos.reshape(-1, 2).to(torch.int32).contiguous()
    return AttentionFn.apply(q_src, kv_src, B, Nq, Nk, H, q_off, k_off, v_off, q32, k32, (base, f0, table), scale)
