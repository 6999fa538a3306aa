"""Hand-scheduled forward / backward of the DUSt3R hot path on top of the C-ABI kernels (ops.py).

Everything here works on token-major 2-D bf16 activations `[rows = B*N, C]`; all matrices come from a
`ParamPack` (bf16 operand copies, fp32 gradient views).  Backward passes are written out explicitly
(no autograd inside): they consume the tensors the forward saved and ACCUMULATE parameter gradients
straight into the pack's flat fp32 gradient buffer (wgrad GEMM = split-K + red.add, bias grads =
column sums, LayerNorm grads = atomics), which is what dp.py all-reduces.

Reference arithmetic being reproduced (file:line under /root/reference/uniception/models):
  encoder block   libs/croco/blocks.py:105-130, :158-161
  decoder block   utils/transformer_blocks.py:208-257, :320-386, :617-647
  encoder         encoders/croco.py:147-182;  decoder  info_sharing/cross_attention_transformer.py:191-275
  linear head     prediction_heads/linear.py:61-84 + adaptors.py:337-342, :1080-1083
Autocast dtype map reproduced by hand: bf16 residual stream / GEMM operands, fp32 accumulation,
fp32 LayerNorm statistics, fp32 softmax, fp32 RoPE angle math, fp32 head output + adaptor.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .params import ParamPack

LN_EPS = 1e-6

# ------------------------------------------------------------------------------------------------
# positions / rope tables (cached per device; index ops, bit-exact)
# ------------------------------------------------------------------------------------------------
_pos_cache: Dict[Tuple, torch.Tensor] = {}
_table_cache: Dict[Tuple, torch.Tensor] = {}


def grid_positions(b: int, h: int, w: int, device, dtype=torch.int32) -> torch.Tensor:
    """[(b*h*w), 2] (y, x), y outer -- libs/croco/patch_embed.py:25-31, keyed by device too."""
    key = (b, h, w, str(device), dtype)
    t = _pos_cache.get(key)
    if t is None:
        ys = torch.arange(h, device=device, dtype=dtype).repeat_interleave(w)
        xs = torch.arange(w, device=device, dtype=dtype).repeat(h)
        t = torch.stack([ys, xs], dim=-1).repeat(b, 1).contiguous()
        _pos_cache[key] = t
    return t


def rope_table(num_pos: int, base: float, f0: float, device) -> torch.Tensor:
    key = (num_pos, float(base), float(f0), str(device))
    t = _table_cache.get(key)
    if t is None:
        t = ops.rope2d_table(num_pos, base, f0, device)
        _table_cache[key] = t
    return t


class Rope:
    """Fused-RoPE context for a token grid: int32 positions [rows,2] + (cos,sin) table.
    `portrait` (per-sample flags, ManyAR patch-embed): those samples use the (w, h) grid (patch_embed.py:120-121)."""

    def __init__(self, b: int, h: int, w: int, base: float, f0: float, device, portrait: Optional[Sequence[bool]] = None):
        if portrait is None or not any(portrait):
            self.pos = grid_positions(b, h, w, device)
        else:
            land, port = grid_positions(1, h, w, device), grid_positions(1, w, h, device)
            self.pos = torch.cat([port if f else land for f in portrait], dim=0).contiguous()
        self.table = rope_table(max(h, w), base, f0, device)


# ------------------------------------------------------------------------------------------------
# linear helpers
# ------------------------------------------------------------------------------------------------
PATCH_EMBED_TMA = os.environ.get("UC_PATCH_EMBED_TMA", "1") != "0"  # 0: uc_patchify + uc_gemm forward (A/B, debugging)


def _empty(rows, cols, like, dtype=torch.bfloat16):
    return torch.empty(rows, cols, dtype=dtype, device=like.device)


def linear_fwd(pk: ParamPack, name: str, x, *, residual=None, rope: Optional[Rope] = None, rope_cols=0, gelu=False,
               out_dtype=torch.bfloat16, w16=None, bias=None):
    w = pk.w16(name + ".weight") if w16 is None else w16
    b = pk.w32(name + ".bias") if bias is None else bias
    out = _empty(x.shape[0], w.shape[0], x, out_dtype)
    if gelu:
        pre = _empty(x.shape[0], w.shape[0], x)
        ops.gemm(x, w, out, bias=b, gelu=True, aux_out=pre)
        return out, pre
    ops.gemm(x, w, out, bias=b, residual=residual,
             positions=rope.pos if rope is not None else None, rope_table=rope.table if rope is not None else None,
             rope_cols=rope_cols)
    return out


class _ColsumSide:
    """Bias-gradient column sums that no producer kernel could fuse (q/k/v projections, proj_embed) are HBM-bound 10-35 us
    kernels; on a side stream they run underneath the tensor-bound wgrad / dgrad GEMMs of the same Linear instead of in
    front of them.  fork: the side stream waits for the caller's stream (dy is complete); join: the caller's stream waits
    for the side stream before `linear_bwd` returns, so dy's block is not recycled early and the bias gradient is final
    before the block is reported to the gradient all-reduce.  UC_COLSUM_STREAM=0 keeps everything on one stream."""

    enabled = os.environ.get("UC_COLSUM_STREAM", "1") != "0"
    _cache: Dict[tuple, "torch.cuda.Stream"] = {}

    @classmethod
    def run(cls, dy, gb):
        """Launch colsum(dy) -> gb; returns the stream to join, or None when it ran on the caller's stream."""
        if not cls.enabled:
            ops.colsum_(dy, gb)
            return None
        cur = torch.cuda.current_stream(dy.device)
        key = (str(dy.device), cur.cuda_stream)
        side = cls._cache.get(key)
        if side is None:
            side = cls._cache[key] = torch.cuda.Stream(dy.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ops.colsum_(dy, gb)
        return side


def linear_bwd(pk: ParamPack, name: str, dy, x_in, *, need_dx=True, gelu_pre=None, w16=None, wgrad=None, bgrad=None,
               train_w=None, bias_done=False, dx_sink=None):
    """dy [rows, out] bf16, x_in [rows, in] bf16.  Accumulates dW, db; returns dx (bf16) or None.
    gelu_pre: dx is additionally multiplied by gelu'(gelu_pre) (the producer of x_in was GELU).
    bias_done: db = colsum(dy) was already accumulated by the kernel that produced dy (uc_layernorm_bwd's dx_colsum or
    uc_gemm's c_colsum); dx_sink: fp32 [in] buffer that receives colsum(dx) from this call's dgrad GEMM epilogue."""
    w = pk.w16(name + ".weight") if w16 is None else w16
    if train_w is None:
        train_w = pk.requires_grad(name + ".weight")
    side = None
    if train_w:
        if not bias_done:
            gb = pk.grad(name + ".bias") if bgrad is None else bgrad
            side = _ColsumSide.run(dy, gb)
        gw = pk.grad(name + ".weight") if wgrad is None else wgrad
        ops.gemm(dy, x_in, gw, a_layout=1, b_layout=1, atomic=True)
    dx = None
    if need_dx:
        dx = _empty(dy.shape[0], w.shape[1], dy)
        if gelu_pre is not None:
            ops.gemm(dy, w, dx, b_layout=1, gelu_bwd=True, aux_in=gelu_pre, c_colsum=dx_sink)
        else:
            ops.gemm(dy, w, dx, b_layout=1, c_colsum=dx_sink)
    if side is not None:
        torch.cuda.current_stream(dy.device).wait_stream(side)
    return dx


def ln_fwd(pk: ParamPack, name: str, x, out_dtype=torch.bfloat16):
    return ops.layernorm_fwd(x, pk.w32(name + ".weight"), pk.w32(name + ".bias"), LN_EPS, out_dtype)


def ln_bwd(pk: ParamPack, name: str, dy, x, mean, rstd, dres=None, colsum=None):
    """colsum: fp32 [C] bias-gradient buffer of the Linear that will consume this call's dx as ITS output gradient
    (see `bias_sink`); the column sums of dx are accumulated there by the same kernel."""
    train = pk.requires_grad(name + ".weight")
    fused_sum = colsum if x.dtype == torch.bfloat16 else None  # the fused column sums live in the bf16 row kernel only
    dx = ops.layernorm_bwd(dy, x, pk.w32(name + ".weight"), mean, rstd,
                           pk.grad(name + ".weight") if train else None, pk.grad(name + ".bias") if train else None, dres=dres,
                           dx_colsum=fused_sum)
    if colsum is not None and fused_sum is None:  # fp32 residual stream (global / alternating transformers): one extra pass
        ops.colsum_(dx, colsum)
    return dx


def has_param(pk: ParamPack, name: str) -> bool:
    return name in pk.params


def ls_name(pk: ParamPack, block_prefix: str, k: int) -> Optional[str]:
    """`{block}.ls{k}.gamma` when the block was built with LayerScale (init_values, utils/transformer_blocks.py:389-412)."""
    n = f"{block_prefix}ls{k}.gamma"
    return n if n in pk.params else None


def out_proj_fwd(pk: ParamPack, name: str, inp, x, ls: Optional[str]):
    """x + [gamma *] Linear(inp): the sub-block's output projection with the residual add.  Returns (new x, z) where z is the
    un-scaled projection (saved for the LayerScale backward) or None."""
    if ls is None:  # the stream keeps its dtype: bf16, or fp32 in the transformers whose reference stream is fp32 under autocast
        return linear_fwd(pk, name, inp, residual=x, out_dtype=x.dtype), None
    z = linear_fwd(pk, name, inp)
    return ops.layerscale_fwd(z, x, pk.w32(ls)), z


def out_proj_grad(pk: ParamPack, dx2, z, ls: Optional[str]):
    """Gradient w.r.t. the un-scaled projection output (LayerScale backward; identity without it)."""
    if ls is None:
        return dx2
    return ops.layerscale_bwd(dx2, z, pk.w32(ls), pk.grad(ls))


def headnorm_fwd(pk: ParamPack, name: str, x, y, rope: Optional[Rope]):
    """qk_norm: y = rope(LayerNorm_64(x)) on every head segment (utils/transformer_blocks.py:222-229, :347-358)."""
    ops.headnorm_fwd(x, y, pk.w32(name + ".weight"), pk.w32(name + ".bias"), LN_EPS,
                     rope.pos if rope is not None else None, rope.table if rope is not None else None)


def headnorm_bwd(pk: ParamPack, name: str, g, x):
    ops.headnorm_bwd(g, x, pk.w32(name + ".weight"), pk.grad(name + ".weight"), pk.grad(name + ".bias"), LN_EPS)


_SINK_SUFFIXES = ("cross_attn.proj", "attn.proj", "mlp.fc2")


def bias_sink(pk: ParamPack, linear_name: Optional[str], like: Optional[torch.Tensor] = None):
    """Gradient buffer of `linear_name`.bias if that Linear trains and the fused column sum applies, else None.
    A block with LayerScale scales the residual-stream gradient before it reaches the projection, so its output
    projections take the un-fused bias-gradient path."""
    if linear_name is None or not pk.requires_grad(linear_name + ".weight"):
        return None
    for suf in _SINK_SUFFIXES:
        if linear_name.endswith(suf):
            if (linear_name[:-len(suf)] + "ls1.gamma") in pk.params:
                return None
            break
    g = pk.grad(linear_name + ".bias")
    if g.shape[0] % 128 != 0 or g.shape[0] > 1024:  # uc_layernorm_bwd's fast path (csrc/elementwise.cu)
        return None
    return g


# ------------------------------------------------------------------------------------------------
# self-attention + MLP sub-blocks (shared by encoder and decoder blocks)
# ------------------------------------------------------------------------------------------------
def attn_scale(softmax_scaling, n_queries: int, head_dim: int = 64) -> float:
    """head_dim^-0.5 times the query multipliers of `use_scalable_softmax` (log N, utils/transformer_blocks.py:231-233,
    :360-362) and `use_entropy_scaling` (sqrt(growth * log N / log base), :235-241, :364-370).  Both only multiply q by a
    scalar that depends on the token count, so they fold into the attention kernels' `scale` (forward, dq and dk alike).
    softmax_scaling: None or (use_scalable_softmax, use_entropy_scaling, base_token_count, growth_factor)."""
    s = head_dim ** -0.5
    if softmax_scaling:
        use_ss, use_es, base, growth = softmax_scaling
        if use_ss:
            s *= math.log(n_queries)
        if use_es:
            s *= math.sqrt(growth * math.log(n_queries) / math.log(base))
    return s



def self_attn_fwd(pk, p, x, B, N, H, rope: Optional[Rope], norm: str, saved: list, scale: float = 0.125, ls: Optional[str] = None):
    """x += [ls *] proj(attn(rope([qk_norm] qkv(LN(x)))))  -- returns the new residual stream."""
    C = H * 64
    h1, mean, rstd = ln_fwd(pk, p + norm, x)
    if has_param(pk, p + "attn.q_norm.weight"):  # qk_norm: raw projection, then LayerNorm_64 + RoPE on the q and k head segments
        qkv = linear_fwd(pk, p + "attn.qkv", h1)
        qk = _empty(qkv.shape[0], 2 * C, qkv)
        headnorm_fwd(pk, p + "attn.q_norm", qkv[:, :C], qk[:, :C], rope)
        headnorm_fwd(pk, p + "attn.k_norm", qkv[:, C:2 * C], qk[:, C:], rope)
        q, k = qk[:, :C], qk[:, C:]
    else:
        qkv = linear_fwd(pk, p + "attn.qkv", h1, rope=rope, rope_cols=2 * C)
        qk, q, k = None, qkv[:, :C], qkv[:, C:2 * C]
    o, lse = ops.attn_fwd(q, k, qkv[:, 2 * C:], B, H, N, N, scale)
    x2, z = out_proj_fwd(pk, p + "attn.proj", o, x, ls)
    saved.append((x, mean, rstd, h1, qkv, o, lse, scale, qk, z))
    return x2


def self_attn_bwd(pk, p, dx2, B, N, H, rope: Optional[Rope], norm: str, saved, bias_done=False, out_sink=None,
                  ls: Optional[str] = None):
    """dx2: gradient w.r.t. the sub-block output (bf16).  Returns gradient w.r.t. its input.
    bias_done: colsum(dx2) already sits in attn.proj.bias.grad; out_sink: bias-gradient buffer fed by the returned dx."""
    x, mean, rstd, h1, qkv, o, lse, scale, qk, z = saved
    C = H * 64
    d_o = linear_bwd(pk, p + "attn.proj", out_proj_grad(pk, dx2, z, ls), o, bias_done=bias_done and ls is None)
    dqkv = torch.empty_like(qkv)
    q, k = (qkv[:, :C], qkv[:, C:2 * C]) if qk is None else (qk[:, :C], qk[:, C:])
    ops.attn_bwd(q, k, qkv[:, 2 * C:], o, d_o, lse, B, H, N, N, scale,
                 dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:],
                 q_positions=rope.pos if rope is not None else None, k_positions=rope.pos if rope is not None else None,
                 rope_table=rope.table if rope is not None else None)
    if qk is not None:
        headnorm_bwd(pk, p + "attn.q_norm", dqkv[:, :C], qkv[:, :C])
        headnorm_bwd(pk, p + "attn.k_norm", dqkv[:, C:2 * C], qkv[:, C:2 * C])
    d_h1 = linear_bwd(pk, p + "attn.qkv", dqkv, h1)
    return ln_bwd(pk, p + norm, d_h1, x, mean, rstd, dres=dx2, colsum=out_sink)


def mlp_fwd(pk, p, x, norm: str, saved: list, ls: Optional[str] = None):
    h, mean, rstd = ln_fwd(pk, p + norm, x)
    act, pre = linear_fwd(pk, p + "mlp.fc1", h, gelu=True)
    x2, z = out_proj_fwd(pk, p + "mlp.fc2", act, x, ls)
    saved.append((x, mean, rstd, h, pre, act, z))
    return x2


def mlp_bwd(pk, p, dx2, norm: str, saved, bias_done=False, out_sink=None, ls: Optional[str] = None):
    x, mean, rstd, h, pre, act, z = saved
    fc1_sink = pk.grad(p + "mlp.fc1.bias") if pk.requires_grad(p + "mlp.fc1.weight") else None
    d_pre = linear_bwd(pk, p + "mlp.fc2", out_proj_grad(pk, dx2, z, ls), act, gelu_pre=pre, bias_done=bias_done and ls is None,
                       dx_sink=fc1_sink)  # fc1.bias.grad = colsum(d_pre)
    d_h = linear_bwd(pk, p + "mlp.fc1", d_pre, h, bias_done=fc1_sink is not None)
    return ln_bwd(pk, p + norm, d_h, x, mean, rstd, dres=dx2, colsum=out_sink)


# ------------------------------------------------------------------------------------------------
# encoder (encoders/croco.py:147-182)
# ------------------------------------------------------------------------------------------------
def encoder_fwd(pk: ParamPack, p: str, img: torch.Tensor, depth: int, heads: int, patch: int, rope_base: Optional[float],
                rope_f0: float = 1.0, take: Sequence[int] = (), norm_intermediate: bool = True,
                portrait: Optional[Sequence[bool]] = None):
    """img fp32 [B,3,H,W] -> (normalised tokens bf16 [B*N, C], intermediates, saved-for-backward).
    portrait: per-sample flags of `ManyAR_PatchEmbed` (libs/croco/patch_embed.py:85-127): flagged samples are patchified
    from the image with its last two axes swapped and take their positions from the (w, h) grid."""
    B, _, Hh, Ww = img.shape
    h, w = Hh // patch, Ww // patch
    N = h * w
    if portrait is not None and not any(portrait):
        portrait = None
    rope = Rope(B, h, w, rope_base, rope_f0, img.device, portrait) if rope_base is not None else None
    wname = p + "patch_embed.proj.weight"
    direct = (portrait is None and PATCH_EMBED_TMA and ops.patch_embed_ok(img, patch, pk.w32(wname).shape[0]))
    if direct:
        # im2col-free: 5-D TMA boxes of the fp32 image feed TF32 MMAs on the fp32 master weight (uc_patch_embed); nothing but the
        # image is kept for the backward pass, which gathers the bf16 columns it needs for the weight gradient itself
        x = ops.patch_embed(img, pk.w32(wname).view(pk.w32(wname).shape[0], -1), pk.w32(p + "patch_embed.proj.bias"), patch)
        saved = {"cols": None, "img": img, "patch": patch, "blocks": [], "B": B, "N": N, "rope": rope, "inter": []}
    elif portrait is None:
        cols = ops.patchify(img, patch)
    else:  # mixed batch: gather each orientation group separately, then restore the sample order (index ops, bit-exact)
        pidx = [i for i, f in enumerate(portrait) if f]
        lidx = [i for i, f in enumerate(portrait) if not f]
        cp = ops.patchify(img[pidx].swapaxes(-1, -2).contiguous(), patch)
        cols = torch.empty(B, N, cp.shape[1], dtype=cp.dtype, device=cp.device)
        cols[pidx] = cp.view(len(pidx), N, -1)
        if lidx:
            cols[lidx] = ops.patchify(img[lidx].contiguous(), patch).view(len(lidx), N, -1)
        cols = cols.view(B * N, -1)
    if not direct:
        wpe = pk.w16(wname)
        if wpe.shape[1] != cols.shape[1]:  # patch sizes whose 3*p*p is not a multiple of 8 (p = 14): zero-padded operand copy
            wpad = torch.zeros(wpe.shape[0], cols.shape[1], dtype=wpe.dtype, device=wpe.device)
            wpad[:, :wpe.shape[1]] = wpe
            wpe = wpad
        x = _empty(cols.shape[0], wpe.shape[0], cols)
        ops.gemm(cols, wpe, x, bias=pk.w32(p + "patch_embed.proj.bias"))
        saved = {"cols": cols, "blocks": [], "B": B, "N": N, "rope": rope, "inter": []}
    inter = []
    for i in range(depth):
        bs: list = []
        bp = f"{p}enc_blocks.{i}."
        x = self_attn_fwd(pk, bp, x, B, N, heads, rope, "norm1", bs)
        x = mlp_fwd(pk, bp, x, "norm2", bs)
        saved["blocks"].append(bs)
        if i in take:
            if norm_intermediate:
                y, m, r = ln_fwd(pk, p + "enc_norm", x)
                saved["inter"].append((i, x, m, r))
                inter.append(y)
            else:
                saved["inter"].append((i, None, None, None))
                inter.append(x)
    y, mean, rstd = ln_fwd(pk, p + "enc_norm", x)
    saved["final"] = (x, mean, rstd)
    return y, inter, saved


def encoder_bwd(pk: ParamPack, p: str, saved, d_out: Optional[torch.Tensor], depth: int, heads: int,
                d_inter: Sequence[Optional[torch.Tensor]] = ()):
    """d_out: gradient w.r.t. the normalised output tokens ([B*N, C] bf16 or fp32), or None."""
    B, N, rope = saved["B"], saved["N"], saved["rope"]
    x, mean, rstd = saved["final"]
    inter_at = {i: (k, xi, m, r) for k, (i, xi, m, r) in enumerate(saved["inter"])}
    # Bias gradients of fc2 / proj are column sums of the residual-stream gradient; the LayerNorm-backward kernel that
    # PRODUCES that gradient accumulates them (`sink`), unless an intermediate-feature tap still adds to it afterwards.
    sink = bias_sink(pk, f"{p}enc_blocks.{depth - 1}.mlp.fc2") if (depth > 0 and (depth - 1) not in inter_at) else None
    dx = ln_bwd(pk, p + "enc_norm", d_out.contiguous(), x, mean, rstd, colsum=sink) if d_out is not None else None
    done = dx is not None and sink is not None
    for i in reversed(range(depth)):
        if i in inter_at:
            k, xi, m, r = inter_at[i]
            g = d_inter[k] if k < len(d_inter) else None
            if g is not None:
                g = g.contiguous()
                if xi is not None:  # normalised intermediate: through the shared enc_norm
                    dx = ln_bwd(pk, p + "enc_norm", g, xi, m, r, dres=dx)
                else:
                    dx = g.to(torch.bfloat16) if dx is None else (dx + g.to(torch.bfloat16))
        if dx is None:
            continue
        bs = saved["blocks"][i]
        bp = f"{p}enc_blocks.{i}."
        sink = bias_sink(pk, bp + "attn.proj")
        dx = mlp_bwd(pk, bp, dx, "norm2", bs[1], bias_done=done, out_sink=sink)
        if i > 0:
            nxt = bias_sink(pk, f"{p}enc_blocks.{i - 1}.mlp.fc2") if (i - 1) not in inter_at else None
        else:
            nxt = bias_sink(pk, p + "patch_embed.proj")
        dx = self_attn_bwd(pk, bp, dx, B, N, heads, rope, "norm1", bs[0], bias_done=sink is not None, out_sink=nxt)
        done = nxt is not None
        pk.notify_done(bp)  # this block's gradients are final -> its all-reduce bucket may start
    if dx is not None and pk.requires_grad(p + "patch_embed.proj.weight"):
        gw = pk.grad(p + "patch_embed.proj.weight")
        cols = saved["cols"]
        if cols is None:  # forward ran im2col-free: gather the bf16 columns now, for the weight gradient only
            cols = ops.patchify(saved["img"], saved["patch"])
        if gw.shape[1] == cols.shape[1]:
            ops.gemm(dx, cols, gw, a_layout=1, b_layout=1, atomic=True)
        else:  # padded pitch (p = 14): accumulate into a padded fp32 scratch, then add the valid columns
            gpad = torch.zeros(gw.shape[0], cols.shape[1], dtype=torch.float32, device=gw.device)
            ops.gemm(dx, cols, gpad, a_layout=1, b_layout=1, atomic=True)
            gw.add_(gpad[:, :gw.shape[1]])
        if not (done and depth > 0):
            ops.colsum_(dx, pk.grad(p + "patch_embed.proj.bias"))
    return None


# ------------------------------------------------------------------------------------------------
# per-view CUDA streams (two-view decoder)
# ------------------------------------------------------------------------------------------------
class ViewStreams:
    """At one decoder depth the two views' blocks only read the PREVIOUS depth's tokens, so their kernel chains are
    independent.  Each chain is ~20 launches of small (m = 8192, n = 768 ...) kernels whose tiles do not fill whole
    waves of the 148 SMs; on two streams the second chain's CTAs take the SMs the first chain's partial waves leave
    idle.  Discipline: `fork()` (view streams wait for the caller's stream) -> per-view work under `on(v)` -> `join()`
    (caller's stream waits for both) once per depth.  Every tensor that crosses streams is either saved for backward
    (long-lived) or consumed before the next fork, so the caching allocator's per-stream pools never recycle a block
    that another stream still reads.  UC_VIEW_STREAMS=0 runs everything on the caller's stream."""

    _cache: Dict[str, List[torch.cuda.Stream]] = {}
    enabled = os.environ.get("UC_VIEW_STREAMS", "1") != "0"

    def __init__(self, nv: int, device):
        self.active = bool(self.enabled and nv == 2 and torch.device(device).type == "cuda")
        self.main = torch.cuda.current_stream(device) if self.active else None
        if self.active:
            key = str(device)
            if key not in self._cache:
                self._cache[key] = [torch.cuda.Stream(device) for _ in range(2)]
            self.streams = self._cache[key]

    def fork(self) -> None:
        if self.active:
            for st in self.streams:
                st.wait_stream(self.main)

    def join(self) -> None:
        if self.active:
            for st in self.streams:
                self.main.wait_stream(st)

    def on(self, v: int):
        return torch.cuda.stream(self.streams[v]) if self.active else contextlib.nullcontext()

    def event(self, v: int):
        """Event recorded on view v's stream (None when inactive)."""
        if not self.active:
            return None
        e = torch.cuda.Event()
        e.record(self.streams[v])
        return e

    def wait(self, v: int, event) -> None:
        if self.active and event is not None:
            self.streams[v].wait_event(event)


# ------------------------------------------------------------------------------------------------
# two-view (N-view) cross-attention decoder (info_sharing/cross_attention_transformer.py:191-275)
# ------------------------------------------------------------------------------------------------
def _cross_fwd(pk, p, x, y, B, Nq, Nk, H, rope_q: Optional[Rope], rope_k: Optional[Rope], saved: list, has_norm_y: bool,
               scale: float = 0.125, ls: Optional[str] = None):
    C = H * 64
    qkn = has_param(pk, p + "cross_attn.q_norm.weight")
    if has_norm_y:
        yn, ymean, yrstd = ln_fwd(pk, p + "norm_y", y)
    else:
        yn, ymean, yrstd = y, None, None
    h2, mean, rstd = ln_fwd(pk, p + "norm2", x)
    q = linear_fwd(pk, p + "cross_attn.projq", h2, rope=None if qkn else rope_q, rope_cols=C)
    # projk | projv are adjacent in the pack -> one GEMM with N = 2C, RoPE on the k half only
    wkv = pk.w16_rows(p + "cross_attn.projk.weight", p + "cross_attn.projv.weight")
    bkv = pk.w32_span(p + "cross_attn.projk.bias", p + "cross_attn.projv.bias")
    kv = linear_fwd(pk, "", yn, rope=None if qkn else rope_k, rope_cols=C, w16=wkv, bias=bkv)
    if qkn:  # q, k: raw projections kept for the backward; normalised + rotated copies feed the attention
        qn, kn = torch.empty_like(q), _empty(kv.shape[0], C, kv)
        headnorm_fwd(pk, p + "cross_attn.q_norm", q, qn, rope_q)
        headnorm_fwd(pk, p + "cross_attn.k_norm", kv[:, :C], kn, rope_k)
    else:
        qn, kn = q, kv[:, :C]
    o, lse = ops.attn_fwd(qn, kn, kv[:, C:], B, H, Nq, Nk, scale)
    x2, z = out_proj_fwd(pk, p + "cross_attn.proj", o, x, ls)
    saved.append((x, mean, rstd, h2, q, kv, o, lse, y, yn, ymean, yrstd, scale, qn if qkn else None, kn if qkn else None, z))
    return x2


def _cross_bwd(pk, p, dx2, B, Nq, Nk, H, rope_q, rope_k, saved, has_norm_y: bool, bias_done=False, out_sink=None,
               ls: Optional[str] = None):
    """Returns (dx, d_yn): gradient w.r.t. the block's own stream and w.r.t. norm_y(y) (bf16)."""
    x, mean, rstd, h2, q, kv, o, lse, y, yn, ymean, yrstd, scale, qn, kn, z = saved
    C = H * 64
    d_o = linear_bwd(pk, p + "cross_attn.proj", out_proj_grad(pk, dx2, z, ls), o, bias_done=bias_done and ls is None)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    ops.attn_bwd(q if qn is None else qn, kv[:, :C] if kn is None else kn, kv[:, C:], o, d_o, lse, B, H, Nq, Nk, scale,
                 dq, dkv[:, :C], dkv[:, C:],
                 q_positions=rope_q.pos if rope_q is not None else None,
                 k_positions=rope_k.pos if rope_k is not None else None,
                 rope_table=rope_q.table if rope_q is not None else None)
    if qn is not None:
        headnorm_bwd(pk, p + "cross_attn.q_norm", dq, q)
        headnorm_bwd(pk, p + "cross_attn.k_norm", dkv[:, :C], kv[:, :C])
    d_h2 = linear_bwd(pk, p + "cross_attn.projq", dq, h2)
    wkv = pk.w16_rows(p + "cross_attn.projk.weight", p + "cross_attn.projv.weight")
    gkv = pk.grad_span(p + "cross_attn.projk.weight", p + "cross_attn.projv.weight").view(2 * C, -1)
    gbkv = pk.grad_span(p + "cross_attn.projk.bias", p + "cross_attn.projv.bias")
    d_yn = linear_bwd(pk, "", dkv, yn, w16=wkv, wgrad=gkv, bgrad=gbkv, train_w=pk.requires_grad(p + "cross_attn.projk.weight"))
    dx = ln_bwd(pk, p + "norm2", d_h2, x, mean, rstd, dres=dx2, colsum=out_sink)
    return dx, d_yn


def decoder_fwd(pk: ParamPack, p: str, toks: List[torch.Tensor], B: int, h: int, w: int, depth: int, heads: int,
                rope_base: Optional[float], rope_f0: float = 1.0, take: Sequence[int] = (), norm_intermediate: bool = True,
                has_proj_embed: bool = True, has_norm_y: bool = True, softmax_scaling=None):
    """toks: per-view bf16 [B*N, C_in].  Returns (per-view normalised bf16 [B*N, dim], intermediates, saved).
    softmax_scaling: see `attn_scale` (the token count is the number of QUERY tokens, N, for self- and cross-attention)."""
    nv = len(toks)
    N = h * w
    dev = toks[0].device
    rope = Rope(B, h, w, rope_base, rope_f0, dev) if rope_base is not None else None
    rope_o = Rope(B * (nv - 1), h, w, rope_base, rope_f0, dev) if (rope_base is not None and nv > 2) else rope
    saved = {"in": toks, "blocks": [], "B": B, "N": N, "nv": nv, "rope": rope, "rope_o": rope_o, "inter": [], "final": []}
    vs = ViewStreams(nv, dev)
    vs.fork()
    xs = []
    for v, t in enumerate(toks):
        with vs.on(v):
            xs.append(linear_fwd(pk, p + "proj_embed", t) if has_proj_embed else t)
    vs.join()
    inter = []
    for k in range(depth):
        new, lvl = [], []
        vs.fork()
        for v in range(nv):
            bp = f"{p}multi_view_branches.{v}.{k}."
            bs: list = []
            with vs.on(v):
                if nv == 2:
                    y = xs[1 - v]
                else:  # other views concatenated along tokens, per batch element
                    y = torch.cat([xs[i].view(B, N, -1) for i in range(nv) if i != v], dim=1).reshape(B * N * (nv - 1), -1)
                sc = attn_scale(softmax_scaling, N)
                x = self_attn_fwd(pk, bp, xs[v], B, N, heads, rope, "norm1", bs, sc, ls_name(pk, bp, 1))
                x = _cross_fwd(pk, bp, x, y, B, N, N * (nv - 1), heads, rope, rope_o, bs, has_norm_y, sc, ls_name(pk, bp, 2))
                x = mlp_fwd(pk, bp, x, "norm3", bs, ls_name(pk, bp, 3))
            new.append(x)
            lvl.append(bs)
        vs.join()
        xs = new
        saved["blocks"].append(lvl)
        if k in take:
            if norm_intermediate:
                outs = []
                for v in range(nv):
                    yv, m, r = ln_fwd(pk, p + "norm", xs[v])
                    outs.append(yv)
                    saved["inter"].append((k, v, xs[v], m, r))
                inter.append(outs)
            else:
                for v in range(nv):
                    saved["inter"].append((k, v, None, None, None))
                inter.append(list(xs))
    outs = []
    for v in range(nv):
        yv, m, r = ln_fwd(pk, p + "norm", xs[v])
        outs.append(yv)
        saved["final"].append((xs[v], m, r))
    return outs, inter, saved


def decoder_bwd(pk: ParamPack, p: str, saved, d_outs: Sequence[Optional[torch.Tensor]], depth: int, heads: int,
                d_inter: Sequence[Sequence[Optional[torch.Tensor]]] = (), has_proj_embed: bool = True, has_norm_y: bool = True,
                need_input_grad: bool = True):
    B, N, nv, rope, rope_o = saved["B"], saved["N"], saved["nv"], saved["rope"], saved["rope_o"]
    inter_levels = sorted({k for (k, _, _, _, _) in saved["inter"]})
    vs = ViewStreams(nv, saved["in"][0].device)

    def stream_sink(v: int, k: int):
        """Bias buffer fed by the FINAL gradient of view v's token stream entering level k from above (k = -1: the
        decoder input): fc2 of block (v, k), or proj_embed; None when an intermediate tap still adds to that gradient."""
        if k >= 0:
            return None if k in inter_levels else bias_sink(pk, f"{p}multi_view_branches.{v}.{k}.mlp.fc2")
        return bias_sink(pk, p + "proj_embed") if has_proj_embed else None

    dxs: List[Optional[torch.Tensor]] = [None] * nv
    done = [False] * nv  # colsum(dxs[v]) already accumulated into its consumer's bias gradient
    for v in range(nv):
        if d_outs[v] is not None:
            x, m, r = saved["final"][v]
            sink = stream_sink(v, depth - 1)
            dxs[v] = ln_bwd(pk, p + "norm", d_outs[v].contiguous(), x, m, r, colsum=sink)
            done[v] = sink is not None
    for k in reversed(range(depth)):
        if k in inter_levels:
            li = inter_levels.index(k)
            for (kk, v, xi, m, r) in saved["inter"]:
                if kk != k:
                    continue
                g = d_inter[li][v] if li < len(d_inter) and d_inter[li] is not None else None
                if g is None:
                    continue
                g = g.contiguous()
                if xi is not None:
                    dxs[v] = ln_bwd(pk, p + "norm", g, xi, m, r, dres=dxs[v])
                else:
                    dxs[v] = g.to(torch.bfloat16) if dxs[v] is None else dxs[v] + g.to(torch.bfloat16)
        lvl = saved["blocks"][k]
        d_own: List[Optional[torch.Tensor]] = [None] * nv
        d_yn: List[Optional[torch.Tensor]] = [None] * nv
        own_done = [False] * nv
        fold_ln = nv == 2 and has_norm_y  # the norm_y backward of the OTHER view produces the final stream gradient
        vs.fork()
        ready = [None] * nv  # event: view v's chain of this depth (d_own[v], d_yn[v]) has been issued
        for v in range(nv):
            if dxs[v] is None:
                continue
            bp = f"{p}multi_view_branches.{v}.{k}."
            bs = lvl[v]
            with vs.on(v):
                s_cross = bias_sink(pk, bp + "cross_attn.proj")
                dx = mlp_bwd(pk, bp, dxs[v], "norm3", bs[2], bias_done=done[v], out_sink=s_cross, ls=ls_name(pk, bp, 3))
                s_attn = bias_sink(pk, bp + "attn.proj")
                dx, d_yn[v] = _cross_bwd(pk, bp, dx, B, N, N * (nv - 1), heads, rope, rope_o, bs[1], has_norm_y,
                                         bias_done=s_cross is not None, out_sink=s_attn, ls=ls_name(pk, bp, 2))
                # if no fold follows for this view (the other view carries no gradient), norm1's backward is final
                s_own = stream_sink(v, k - 1) if (nv == 2 and dxs[1 - v] is None) else None
                d_own[v] = self_attn_bwd(pk, bp, dx, B, N, heads, rope, "norm1", bs[0], bias_done=s_attn is not None, out_sink=s_own,
                                         ls=ls_name(pk, bp, 1))
            own_done[v] = s_own is not None
            ready[v] = vs.event(v)
        # fold the cross-view gradients: d tokens_v(k-1) = d_own[v] + sum_{u != v} norm_y_u'(d_yn[u])|_v
        new: List[Optional[torch.Tensor]] = list(d_own)
        new_done = list(own_done)
        finished: List[str] = []
        for u in range(nv):
            bp = f"{p}multi_view_branches.{u}.{k}."
            if d_yn[u] is None:
                if dxs[u] is not None:
                    finished.append(bp)
                continue
            y, yn, ymean, yrstd = lvl[u][1][8], lvl[u][1][9], lvl[u][1][10], lvl[u][1][11]
            if nv == 2:
                v = 1 - u
                vs.wait(v, ready[u])  # d_yn[u] comes from the other view's stream
                with vs.on(v):
                    if has_norm_y:
                        sink = stream_sink(v, k - 1) if fold_ln else None
                        new[v] = ln_bwd(pk, bp + "norm_y", d_yn[u], y, ymean, yrstd, dres=new[v], colsum=sink)
                        new_done[v] = sink is not None
                    else:
                        new[v] = d_yn[u] if new[v] is None else new[v] + d_yn[u]
                        new_done[v] = False
            else:
                dy = ln_bwd(pk, bp + "norm_y", d_yn[u], y, ymean, yrstd) if has_norm_y else d_yn[u]
                parts = dy.view(B, nv - 1, N, -1)
                others = [i for i in range(nv) if i != u]
                for j, v in enumerate(others):
                    g = parts[:, j].reshape(B * N, -1)
                    new[v] = g.contiguous() if new[v] is None else new[v] + g
                    new_done[v] = False
            finished.append(bp)
        vs.join()
        for bp in finished:
            pk.notify_done(bp)  # this block's gradients (incl. its norm_y, written from the other view's stream) are final
        dxs, done = new, new_done
    d_in: List[Optional[torch.Tensor]] = [None] * nv
    for v in range(nv):
        if dxs[v] is None:
            continue
        if has_proj_embed:
            d_in[v] = linear_bwd(pk, p + "proj_embed", dxs[v], saved["in"][v], need_dx=need_input_grad, bias_done=done[v])
        else:
            d_in[v] = dxs[v]
    return d_in


# ------------------------------------------------------------------------------------------------
# global / alternating self-attention over all views (info_sharing/global_attention_transformer.py:224-462,
# alternating_attention_transformer.py:397-442): the encoder block arithmetic on the concatenated token set
# ------------------------------------------------------------------------------------------------
def _mv_block_fwd(pk: ParamPack, p: str, i: int, x, Bb: int, Nn: int, heads: int, rope, softmax_scaling):
    """One SelfAttentionBlock on a [Bb*Nn, dim] token buffer -> (new stream, tensors saved for its backward)."""
    bs: list = []
    bp = f"{p}self_attention_blocks.{i}."
    x = self_attn_fwd(pk, bp, x, Bb, Nn, heads, rope, "norm1", bs, attn_scale(softmax_scaling, Nn), ls_name(pk, bp, 1))
    x = mlp_fwd(pk, bp, x, "norm2", bs, ls_name(pk, bp, 2))
    return x, bs


def mv_self_attn_fwd(pk: ParamPack, p: str, x_in: torch.Tensor, B: int, nv: int, n_view: int, n_extra: int, h: int, w: int,
                     depth: int, heads: int, rope_base: Optional[float], rope_f0: float, alternating: bool,
                     view_pe: Optional[torch.Tensor], has_proj_embed: bool, softmax_scaling=None, take: Sequence[int] = (),
                     norm_intermediate: bool = True, recompute: bool = False):
    """x_in: bf16 [B*L, C_in], rows ordered (batch, [view, token], extra) with L = nv*n_view + n_extra: per view its h*w patch
    tokens followed by its per-view additional tokens (n_view of them in total), then n_extra global additional tokens
    (global_attention_transformer.py:266-333).  view_pe: fp32 [V, dim] view-index encodings added after proj_embed to the view
    tokens only (:365-400), or None.  With n_extra == 0 the SAME buffer is a [B, V*n_view] sequence set for the global layers and
    a [B*V, n_view] one for the frame-level layers of the alternating variant: no data movement between the two; with global
    additional tokens the frame layers run on a gathered copy of the view rows and the extra rows skip the block
    (alternating_attention_transformer.py:402-447).
    take / norm_intermediate: intermediate-feature-returner variants (global_attention_transformer.py:766-774).
    recompute: activation checkpointing per block (`gradient_checkpointing=True`: the reference wraps every block in
    torch.utils.checkpoint, info_sharing/base.py:59-71, alternating_attention_transformer.py:178-180) -- only each block's input
    is kept and the backward re-runs the block's forward kernels first.
    Returns (final normalised tokens [B*L, dim], [intermediate [B*L, dim] per taken depth], saved)."""
    dev = x_in.device
    L = nv * n_view + n_extra
    rope = Rope(B * nv, h, w, rope_base, rope_f0, dev) if rope_base is not None else None
    assert rope is None or (n_extra == 0 and n_view == h * w), "RoPE is not defined for additional tokens"
    # Stream dtype.  The reference adds the fp32 view-encoding table to the (bf16 under autocast) projection, which promotes its
    # residual stream to fp32 for the rest of the transformer (global_attention_transformer.py:338-351: `x + pe`, then every
    # `x + attn(...)` stays fp32).  A bf16 stream here measured 1.8x the reference's own autocast error at 12 blocks / 2048
    # tokens; so with a view encoding the stream is fp32 (fp32 residual in / fp32 out of the two projection GEMMs per block,
    # LayerNorm reads fp32).  Blocks with LayerScale keep the bf16 stream (uc_layerscale_* is bf16).
    fp32_stream = view_pe is not None and not has_param(pk, f"{p}self_attention_blocks.0.ls1.gamma")
    sdt = torch.float32 if fp32_stream else torch.bfloat16
    pe = None
    if view_pe is not None:  # [V, dim] -> one row per token (zeros for the global extras), in the stream's dtype
        rows = view_pe.to(sdt).repeat_interleave(n_view, dim=0)
        if n_extra:
            rows = torch.cat([rows, torch.zeros(n_extra, rows.shape[1], dtype=rows.dtype, device=dev)], dim=0)
        pe = rows.repeat(B, 1).contiguous()
    if has_proj_embed:
        x = linear_fwd(pk, p + "proj_embed", x_in, residual=pe, out_dtype=sdt)  # the view encoding rides the GEMM's residual epilogue
    elif pe is None:
        x = x_in
    elif fp32_stream:
        x = x_in.float() + pe
    else:
        x = ops.elementwise(0, x_in.contiguous(), pe)
    saved = {"x_in": x_in, "blocks": [], "B": B, "L": L, "nv": nv, "n_view": n_view, "n_extra": n_extra, "rope": rope, "inter": [],
             "softmax_scaling": softmax_scaling}
    nvt = nv * n_view
    inter = []
    for i in range(depth):
        frame = alternating and i % 2 == 1
        if frame and n_extra:
            x3 = x.view(B, L, -1)
            blk_in = x3[:, :nvt].reshape(B * nvt, -1)
            xv, bs = _mv_block_fwd(pk, p, i, blk_in, B * nv, n_view, heads, rope, softmax_scaling)
            x = torch.cat([xv.view(B, nvt, -1), x3[:, nvt:]], dim=1).reshape(B * L, -1)
        else:
            Bb, Nn = (B * nv, n_view) if frame else (B, L)
            blk_in = x
            x, bs = _mv_block_fwd(pk, p, i, blk_in, Bb, Nn, heads, rope, softmax_scaling)
        saved["blocks"].append((blk_in, None) if recompute else (None, bs))
        if i in take:
            if norm_intermediate:
                yi, m, r = ln_fwd(pk, p + "norm", x)
                saved["inter"].append((i, x, m, r))
                inter.append(yi)
            else:
                saved["inter"].append((i, None, None, None))
                inter.append(x)
    y, mean, rstd = ln_fwd(pk, p + "norm", x)
    saved["final"] = (x, mean, rstd)
    return y, inter, saved


def mv_self_attn_bwd(pk: ParamPack, p: str, saved, d_out: Optional[torch.Tensor], depth: int, heads: int,
                     alternating: bool, has_proj_embed: bool, need_input_grad: bool = True,
                     d_inter: Sequence[Optional[torch.Tensor]] = ()):
    """d_out / d_inter[k]: bf16 [B*L, dim] gradients of the final / tapped outputs (or None).  Returns d x_in or None."""
    B, L, nv, n_view, n_extra, rope = saved["B"], saved["L"], saved["nv"], saved["n_view"], saved["n_extra"], saved["rope"]
    x, mean, rstd = saved["final"]
    nvt = nv * n_view
    # fused bias-gradient sinks hand colsum(dx) of one block to the next; with global extras the frame-level blocks see
    # only the view rows, so the cross-block sinks would sum the wrong row set
    xsink = not (alternating and n_extra)
    inter_at = {i: (k, xi, m, r) for k, (i, xi, m, r) in enumerate(saved["inter"])}
    sink = bias_sink(pk, f"{p}self_attention_blocks.{depth - 1}.mlp.fc2") if (depth > 0 and xsink and (depth - 1) not in inter_at) else None
    dx = ln_bwd(pk, p + "norm", d_out.contiguous(), x, mean, rstd, colsum=sink) if d_out is not None else None
    done = dx is not None and sink is not None
    for i in reversed(range(depth)):
        if i in inter_at:
            k, xi, m, r = inter_at[i]
            g = d_inter[k] if k < len(d_inter) else None
            if g is not None:
                g = g.contiguous()
                if xi is not None:  # normalised intermediate: through the shared final norm
                    dx = ln_bwd(pk, p + "norm", g, xi, m, r, dres=dx)
                else:
                    dx = g.to(torch.bfloat16) if dx is None else ops.elementwise(0, dx, g.to(torch.bfloat16))
        if dx is None:
            continue
        frame = alternating and i % 2 == 1
        blk_in, bs = saved["blocks"][i]
        if bs is None:  # activation checkpointing: regenerate the block's saved tensors from its input
            Bb, Nn = (B * nv, n_view) if frame else (B, L)
            _, bs = _mv_block_fwd(pk, p, i, blk_in, Bb, Nn, heads, rope, saved["softmax_scaling"])
            saved["blocks"][i] = None
        bp = f"{p}self_attention_blocks.{i}."
        sink = bias_sink(pk, bp + "attn.proj")
        if i > 0:
            nxt = bias_sink(pk, f"{p}self_attention_blocks.{i - 1}.mlp.fc2") if (xsink and (i - 1) not in inter_at) else None
        else:
            nxt = bias_sink(pk, p + "proj_embed") if (has_proj_embed and xsink) else None
        if frame and n_extra:
            d3 = dx.view(B, L, -1)
            dv = d3[:, :nvt].reshape(B * nvt, -1)
            dv = mlp_bwd(pk, bp, dv, "norm2", bs[1], bias_done=False, out_sink=sink, ls=ls_name(pk, bp, 2))
            dv = self_attn_bwd(pk, bp, dv, B * nv, n_view, heads, rope, "norm1", bs[0], bias_done=sink is not None, out_sink=None,
                               ls=ls_name(pk, bp, 1))
            dx = torch.cat([dv.view(B, nvt, -1), d3[:, nvt:]], dim=1).reshape(B * L, -1)
        else:
            Bb, Nn = (B * nv, n_view) if frame else (B, L)
            dx = mlp_bwd(pk, bp, dx, "norm2", bs[1], bias_done=done, out_sink=sink, ls=ls_name(pk, bp, 2))
            dx = self_attn_bwd(pk, bp, dx, Bb, Nn, heads, rope, "norm1", bs[0], bias_done=sink is not None, out_sink=nxt,
                               ls=ls_name(pk, bp, 1))
        done = nxt is not None
        pk.notify_done(bp)
    if dx is None:
        return None
    if has_proj_embed:  # the view encoding is a constant: its gradient is dropped
        return linear_bwd(pk, p + "proj_embed", dx, saved["x_in"], need_dx=need_input_grad, bias_done=done and depth > 0)
    return dx


# ------------------------------------------------------------------------------------------------
# linear head + pixel-shuffle + pointmap/confidence adaptor
# ------------------------------------------------------------------------------------------------
def linear_head_fwd(pk: ParamPack, p: str, tok: torch.Tensor, B: int, h: int, w: int, patch: int, conf_min: float, conf_max: float):
    """tok bf16 [B*h*w, C] -> (pts [B,H,W,3], conf [B,H,W,1]) fp32."""
    wt = pk.w16(p + "linear.weight")
    y = _empty(tok.shape[0], wt.shape[0], tok, torch.float32)
    ops.gemm(tok, wt, y, bias=pk.w32(p + "linear.bias"))
    pts, conf = ops.head_post_fwd(y, B, h, w, patch, conf_min, conf_max)
    return pts, conf, (tok, y)


def linear_head_bwd(pk: ParamPack, p: str, saved, dpts, dconf, B, h, w, patch, conf_min, conf_max, need_dx=True):
    tok, y = saved
    dy = ops.head_post_bwd(y, dpts, dconf, B, h, w, patch, conf_min, conf_max)
    dx = linear_bwd(pk, p + "linear", dy, tok, need_dx=need_dx)
    pk.notify_done(p)
    return dx
