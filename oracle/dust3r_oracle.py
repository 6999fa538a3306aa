"""TEST INFRASTRUCTURE ONLY -- CPU/fp32 restatement of the reference's DUSt3R two-view path.

This is the *checker* for the CUDA product in `uniception_b200/`; nothing in the product
path may import it (only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs do).

Parity status: PINNED by differential tests against the real reference
(castacks/UniCeption @ 802ebc17, imported from /root/reference in the build container by
`oracle/make_golden.py`); the resulting vectors are committed under `tests/golden/`.  The
reference's only known-answer test (examples/models/dust3r/dust3r.py:198-230,
`03_head_output.npz`) ships no fixtures, so its *metric form* (max-abs + relative L2) is
what `parity()` below implements.

It is a functional restatement: every function takes plain tensors and a flat
`state_dict` with the reference's key names, and cites the reference
file:line whose arithmetic it follows.  All math runs in whatever dtype the inputs carry
(fp32 for the goldens, fp64 for tighter checks).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# parity metric -- form taken from examples/models/dust3r/dust3r.py:223-230
# --------------------------------------------------------------------------------------
def parity(x: Tensor, ref: Tensor) -> Tuple[float, float]:
    """(max-abs error, relative L2 error ||x-ref||_2 / ||ref||_2)."""
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    diff = x - ref
    denom = float(ref.norm())
    return float(diff.abs().max()) if diff.numel() else 0.0, float(diff.norm()) / (denom if denom > 0 else 1.0)


# --------------------------------------------------------------------------------------
# integer / index ops (bit-exact)
# --------------------------------------------------------------------------------------
def patch_positions(b: int, h: int, w: int, device="cpu") -> Tensor:
    """(y, x) integer grid, row-major with y outer.  libs/croco/patch_embed.py:25-31,
    utils/positional_encoding.py:8-23 (cartesian_prod(arange(h), arange(w)))."""
    ys = torch.arange(h, device=device).repeat_interleave(w)
    xs = torch.arange(w, device=device).repeat(h)
    return torch.stack([ys, xs], dim=-1).view(1, h * w, 2).expand(b, -1, -1).clone()


def feature_take_indices(num_features: int, indices=None) -> Tuple[List[int], int]:
    """utils/intermediate_feature_return.py:47-85: None -> all, int n -> last n, list (negatives ok)."""
    if indices is None:
        indices = num_features
    if isinstance(indices, int):
        assert 0 < indices <= num_features
        take = [num_features - indices + i for i in range(indices)]
    else:
        take = []
        for i in indices:
            idx = num_features + i if i < 0 else i
            assert 0 <= idx < num_features
            take.append(idx)
    return take, max(take)


def is_symmetrized(inst1: Sequence, inst2: Sequence) -> bool:
    """factory/dust3r.py:21-30."""
    if len(inst1) == len(inst2) and len(inst1) == 1:
        return False
    ok = True
    for i in range(0, len(inst1), 2):
        ok = ok and (inst1[i] == inst2[i + 1]) and (inst1[i + 1] == inst2[i])
    return ok


def interleave(t1: Tensor, t2: Tensor) -> Tuple[Tensor, Tensor]:
    """factory/dust3r.py:33-37."""
    r1 = torch.stack((t1, t2), dim=1).flatten(0, 1)
    r2 = torch.stack((t2, t1), dim=1).flatten(0, 1)
    return r1, r2


def pixel_shuffle(x: Tensor, p: int) -> Tensor:
    """out[b,c,p*h+i,p*w+j] = in[b, c*p*p + i*p + j, h, w]  (prediction_heads/linear.py:81-82)."""
    b, cpp, h, w = x.shape
    c = cpp // (p * p)
    return x.view(b, c, p, p, h, w).permute(0, 1, 4, 2, 5, 3).reshape(b, c, h * p, w * p)


# --------------------------------------------------------------------------------------
# float ops
# --------------------------------------------------------------------------------------
# The restatement above/below spells every op out (mean / var / erf / softmax) so that it can be read against the reference
# line by line.  The reference ITSELF calls the fused torch functionals -- nn.LayerNorm -> F.layer_norm, nn.GELU -> F.gelu,
# F.scaled_dot_product_attention by default (utils/config.py:13-17, libs/croco/blocks.py:123-126,
# utils/transformer_blocks.py:244, :373) and, when its CUDA RoPE extension is not built, the PyTorch RoPE fallback with an
# activation-dtype angle table and a host sync per call (libs/croco/pos_embed.py:116-155).  In fp32 the two spellings agree
# to rounding (tests/test_oracle_golden.py); under torch.autocast they do NOT (autocast runs F.layer_norm / softmax in fp32 but
# leaves a hand-written mean/var in bf16), so everything that measures "the reference under bf16 autocast" -- the parity
# yardstick of the GPU tests and bench.py's gpu_eager_baseline -- runs inside `reference_functionals()`.
_FUNCTIONAL = False
_ROPE_CACHE: Dict = {}


class reference_functionals:
    """Context manager: route layer_norm / gelu / sdpa / rope2d through the torch calls the reference makes."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled

    def __enter__(self):
        global _FUNCTIONAL
        self.prev, _FUNCTIONAL = _FUNCTIONAL, self.enabled
        return self

    def __exit__(self, *exc):
        global _FUNCTIONAL
        _FUNCTIONAL = self.prev
        return False


def _rope2d_reference_fallback(tokens: Tensor, positions: Tensor, base: float, f0: float) -> Tensor:
    """libs/croco/pos_embed.py:116-155 as written: cos/sin table cached per (D, seq_len, device, dtype), angles rounded to
    the activation dtype BEFORE cos/sin (:120-123), `int(positions.max()) + 1` = one device->host sync per call (:149)."""
    D = tokens.size(3) // 2
    seq_len = int(positions.max()) + 1
    key = (D, seq_len, str(tokens.device), tokens.dtype, float(base), float(f0))
    if key not in _ROPE_CACHE:
        inv_freq = f0 / (base ** (torch.arange(0, D, 2).float().to(tokens.device) / D))
        t = torch.arange(seq_len, device=tokens.device, dtype=inv_freq.dtype)
        freqs = torch.einsum("i,j->ij", t, inv_freq).to(tokens.dtype)
        freqs = torch.cat((freqs, freqs), dim=-1)
        _ROPE_CACHE[key] = (freqs.cos(), freqs.sin())
    cos, sin = _ROPE_CACHE[key]

    def rope1d(tok, pos1d):
        c = F.embedding(pos1d, cos)[:, None, :, :]
        s_ = F.embedding(pos1d, sin)[:, None, :, :]
        x1, x2 = tok[..., : tok.shape[-1] // 2], tok[..., tok.shape[-1] // 2:]
        return (tok * c) + (torch.cat((-x2, x1), dim=-1) * s_)

    y, x = tokens.chunk(2, dim=-1)
    return torch.cat((rope1d(y, positions[:, :, 0]), rope1d(x, positions[:, :, 1])), dim=-1)


def rope2d(tokens: Tensor, positions: Tensor, base: float = 100.0, f0: float = 1.0) -> Tensor:
    """2-D rotary embedding on tokens [B,H,N,D], positions [B,N,2] (y,x) integers.

    Formula of the native kernel, libs/croco/curope/kernels.cu:39-80 / curope.cpp:21-41
    (identical to the PyTorch fallback libs/croco/pos_embed.py:116-155 in fp32):
    D = 4Q; for half X in {0:y, 1:x}, i < Q: theta = pos[b,n,X] * f0 / base**(i/Q);
    (u, v) = (t[2QX+i], t[2QX+Q+i]) -> (u cos - v sin, v cos + u sin).
    Backward is the same map with f0 -> -f0 (curope2d.py:24-28)."""
    if _FUNCTIONAL:
        return _rope2d_reference_fallback(tokens, positions.long(), base, f0)
    B, H, N, D = tokens.shape
    assert D % 4 == 0
    Q = D // 4
    i = torch.arange(Q, device=tokens.device, dtype=torch.float32)
    inv_freq = (f0 / torch.pow(torch.tensor(base, dtype=torch.float32, device=tokens.device), i / Q))
    ang = positions.to(torch.float32)[..., None] * inv_freq  # [B,N,2,Q] fp32 angle math
    cos = ang.cos().to(tokens.dtype)[:, None]  # [B,1,N,2,Q]
    sin = ang.sin().to(tokens.dtype)[:, None]
    t = tokens.reshape(B, H, N, 2, 2, Q)
    u, v = t[..., 0, :], t[..., 1, :]
    out = torch.stack((u * cos - v * sin, v * cos + u * sin), dim=-2)
    return out.reshape(B, H, N, D)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """nn.LayerNorm(C, eps=1e-6), biased variance (encoders/croco.py:32)."""
    if _FUNCTIONAL:
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() exact-erf form (libs/croco/blocks.py:67)."""
    if _FUNCTIONAL:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    if _FUNCTIONAL:  # nn.Linear: under autocast the bias add and the output are in the autocast dtype too
        return F.linear(x, w, b)
    y = x @ w.t()
    return y if b is None else y + b


def sdpa(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """softmax(q k^T d^-0.5) v per (b,h); the reference's naive path, libs/croco/blocks.py:117-120."""
    if _FUNCTIONAL:
        return F.scaled_dot_product_attention(q, k, v, scale=q.shape[-1] ** -0.5)
    s = (q @ k.transpose(-2, -1)) * (q.shape[-1] ** -0.5)
    return s.softmax(dim=-1) @ v


def mlp(sd: SD, p: str, x: Tensor) -> Tensor:
    """fc2(GELU(fc1 x)), libs/croco/blocks.py:80-86, utils/transformer_blocks.py:82-89."""
    return linear(gelu_erf(linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"])), sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def softmax_q_multiplier(n_tokens: int, softmax_scaling=None) -> float:
    """Query multipliers of `use_scalable_softmax` (log N) and `use_entropy_scaling` (sqrt(growth log N / log base)),
    utils/transformer_blocks.py:231-241, :360-370.  softmax_scaling = (use_ss, use_es, base_token_count, growth) or None."""
    m = 1.0
    if softmax_scaling:
        use_ss, use_es, base_cnt, growth = softmax_scaling
        if use_ss:
            m *= math.log(n_tokens)
        if use_es:
            m *= math.sqrt(growth * math.log(n_tokens) / math.log(base_cnt))
    return m


def qk_norm(sd: SD, p: str, q: Tensor, k: Tensor) -> Tuple[Tensor, Tensor]:
    """`q, k = self.q_norm(q), self.k_norm(k)` with norm_layer(head_dim) on [B,H,N,d] (utils/transformer_blocks.py:199-200,
    :222, :306-307, :347); Identity when the block was built with qk_norm=False (no `q_norm.*` keys)."""
    if (p + "q_norm.weight") not in sd:
        return q, k
    return (layer_norm(q, sd[p + "q_norm.weight"], sd[p + "q_norm.bias"]),
            layer_norm(k, sd[p + "k_norm.weight"], sd[p + "k_norm.bias"]))


def layer_scale(sd: SD, name: str, t: Tensor) -> Tensor:
    """LayerScale (utils/transformer_blocks.py:389-412): t * gamma; Identity when the block has no `ls*.gamma` (init_values=None)."""
    return t * sd[name] if name in sd else t


def self_attention(sd: SD, p: str, x: Tensor, pos: Optional[Tensor], heads: int, base: float, softmax_scaling=None) -> Tensor:
    """libs/croco/blocks.py:105-130 == utils/transformer_blocks.py:208-257 (optional qk_norm and softmax scaling)."""
    B, N, _ = x.shape
    C = sd[p + "qkv.weight"].shape[0] // 3  # == dim, or latent_attn_dim (utils/transformer_blocks.py:178-199)
    d = C // heads
    qkv = linear(x, sd[p + "qkv.weight"], sd.get(p + "qkv.bias")).view(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q, k = qk_norm(sd, p, q, k)
    if pos is not None:
        q, k = rope2d(q, pos, base), rope2d(k, pos, base)
    q = q * softmax_q_multiplier(N, softmax_scaling)
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, N, C)
    return linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])


def cross_attention(sd: SD, p: str, xq: Tensor, y: Tensor, qpos, kpos, heads: int, base: float, softmax_scaling=None) -> Tensor:
    """utils/transformer_blocks.py:320-386."""
    B, Nq, C = xq.shape
    Nk = y.shape[1]
    d = C // heads
    q = linear(xq, sd[p + "projq.weight"], sd.get(p + "projq.bias")).view(B, Nq, heads, d).permute(0, 2, 1, 3)
    k = linear(y, sd[p + "projk.weight"], sd.get(p + "projk.bias")).view(B, Nk, heads, d).permute(0, 2, 1, 3)
    v = linear(y, sd[p + "projv.weight"], sd.get(p + "projv.bias")).view(B, Nk, heads, d).permute(0, 2, 1, 3)
    q, k = qk_norm(sd, p, q, k)
    if qpos is not None:
        q, k = rope2d(q, qpos, base), rope2d(k, kpos, base)
    q = q * softmax_q_multiplier(Nq, softmax_scaling)
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, Nq, C)
    return linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])


def encoder_block(sd: SD, p: str, x: Tensor, pos: Tensor, heads: int, base: float, softmax_scaling=None) -> Tensor:
    """libs/croco/blocks.py:158-161 == SelfAttentionBlock (utils/transformer_blocks.py:497-499; LayerScale when `ls*.gamma` exist)."""
    x = x + layer_scale(sd, p + "ls1.gamma", self_attention(sd, p + "attn.", layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"]),
                                                           pos, heads, base, softmax_scaling))
    x = x + layer_scale(sd, p + "ls2.gamma", mlp(sd, p + "mlp.", layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])))
    return x


def decoder_block(sd: SD, p: str, x: Tensor, y: Tensor, xpos, ypos, heads: int, base: float, softmax_scaling=None) -> Tensor:
    """utils/transformer_blocks.py:643-646 (DropPath is Identity; LayerScale when `ls*.gamma` exist, Identity for DUSt3R)."""
    x = x + layer_scale(sd, p + "ls1.gamma", self_attention(sd, p + "attn.", layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"]),
                                                           xpos, heads, base, softmax_scaling))
    y_ = layer_norm(y, sd[p + "norm_y.weight"], sd[p + "norm_y.bias"])
    x = x + layer_scale(sd, p + "ls2.gamma", cross_attention(
        sd, p + "cross_attn.", layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), y_, xpos, ypos, heads, base,
        softmax_scaling,
    ))
    x = x + layer_scale(sd, p + "ls3.gamma", mlp(sd, p + "mlp.", layer_norm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])))
    return x


# --------------------------------------------------------------------------------------
# DiffAttention family (utils/transformer_blocks.py:658-1031, info_sharing/diff_cross_attention_transformer.py:22-588)
# --------------------------------------------------------------------------------------
def lambda_init(depth: int) -> float:
    """utils/transformer_blocks.py:682-683."""
    return 0.8 - 0.6 * math.exp(-0.3 * depth)


def _diff_combine(sd: SD, p: str, attn1: Tensor, attn2: Tensor, depth: int) -> Tensor:
    """attn1 - lambda_full * attn2 -> RMSNorm(2 d, eps 1e-5, affine) -> x (1 - lambda_init)
    (utils/transformer_blocks.py:779-787 / :924-931; RMSNorm :658-679 normalises in fp32 and casts back)."""
    l1 = torch.exp(torch.sum(sd[p + "lambda_q1"] * sd[p + "lambda_k1"], dim=-1).float()).type_as(attn1)
    l2 = torch.exp(torch.sum(sd[p + "lambda_q2"] * sd[p + "lambda_k2"], dim=-1).float()).type_as(attn1)
    li = lambda_init(depth)
    a = attn1 - (l1 - l2 + li) * attn2
    af = a.float()
    a = (af * torch.rsqrt(af.pow(2).mean(-1, keepdim=True) + 1e-5)).type_as(a) * sd[p + "subln.weight"]
    return a * (1 - li)


def diff_attention(sd: SD, p: str, x: Tensor, pos: Optional[Tensor], heads: int, depth: int, base: float) -> Tensor:
    """`DiffAttention.forward` (utils/transformer_blocks.py:743-795).  heads = the layer's num_heads H: q / k are split into
    2 H heads of d = C / H / 2, v into H heads of 2 d; the [B, H, N, 2 d] result is reshaped to [B, N, C] WITHOUT a
    head / token transpose (:789), as the reference does."""
    B, N, C = x.shape
    d = C // heads // 2
    qkv = linear(x, sd[p + "qkv.weight"], sd.get(p + "qkv.bias")).reshape(B, N, 3, heads, 2 * d)
    q, k, v = torch.chunk(qkv, 3, dim=2)
    q = q.reshape(B, N, 2 * heads, d).permute(0, 2, 1, 3)
    k = k.reshape(B, N, 2 * heads, d).permute(0, 2, 1, 3)
    v = v.reshape(B, N, heads, 2 * d).permute(0, 2, 1, 3)
    q, k = qk_norm(sd, p, q, k)
    if pos is not None:
        q, k = rope2d(q, pos, base), rope2d(k, pos, base)
    q1, q2 = q.chunk(2, dim=1)
    k1, k2 = k.chunk(2, dim=1)
    scale = d ** -0.5
    a1 = ((q1 @ k1.transpose(-2, -1)) * scale).softmax(dim=-1) @ v
    a2 = ((q2 @ k2.transpose(-2, -1)) * scale).softmax(dim=-1) @ v
    a = _diff_combine(sd, p, a1, a2, depth).reshape(B, N, heads * 2 * d)
    return linear(a, sd[p + "proj.weight"], sd[p + "proj.bias"])


def diff_cross_attention(sd: SD, p: str, xq: Tensor, y: Tensor, qpos, kpos, heads: int, depth: int, base: float) -> Tensor:
    """`DiffCrossAttention.forward` (utils/transformer_blocks.py:881-938): as above with separate q / k / v projections and
    the [B, H, Nq, 2 d] -> [B, Nq, H, 2 d] transpose before the combination (:921-922)."""
    B, Nq, C = xq.shape
    Nk = y.shape[1]
    d = C // heads // 2
    q = linear(xq, sd[p + "projq.weight"], sd.get(p + "projq.bias")).reshape(B, Nq, 2 * heads, d).permute(0, 2, 1, 3)
    k = linear(y, sd[p + "projk.weight"], sd.get(p + "projk.bias")).reshape(B, Nk, 2 * heads, d).permute(0, 2, 1, 3)
    v = linear(y, sd[p + "projv.weight"], sd.get(p + "projv.bias")).reshape(B, Nk, heads, 2 * d).permute(0, 2, 1, 3)
    q, k = qk_norm(sd, p, q, k)
    if qpos is not None:
        q, k = rope2d(q, qpos, base), rope2d(k, kpos, base)
    q1, q2 = q.chunk(2, dim=1)
    k1, k2 = k.chunk(2, dim=1)
    scale = d ** -0.5
    a1 = (((q1 @ k1.transpose(-2, -1)) * scale).softmax(dim=-1) @ v).transpose(1, 2)
    a2 = (((q2 @ k2.transpose(-2, -1)) * scale).softmax(dim=-1) @ v).transpose(1, 2)
    a = _diff_combine(sd, p, a1, a2, depth).reshape(B, Nq, heads * 2 * d)
    return linear(a, sd[p + "proj.weight"], sd[p + "proj.bias"])


def diff_decoder_block(sd: SD, p: str, x: Tensor, y: Tensor, xpos, ypos, heads: int, depth: int, base: float) -> Tensor:
    """`DiffCrossAttentionBlock` (utils/transformer_blocks.py:989-1031 on CrossAttentionBlock.forward :620-647): plain
    self-attention with `heads` heads (head_dim = C / heads), differential cross-attention, MLP."""
    x = x + layer_scale(sd, p + "ls1.gamma", self_attention(sd, p + "attn.", layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"]),
                                                           xpos, heads, base))
    y_ = layer_norm(y, sd[p + "norm_y.weight"], sd[p + "norm_y.bias"]) if (p + "norm_y.weight") in sd else y
    x = x + layer_scale(sd, p + "ls2.gamma", diff_cross_attention(
        sd, p + "cross_attn.", layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), y_, xpos, ypos, heads, depth, base))
    x = x + layer_scale(sd, p + "ls3.gamma", mlp(sd, p + "mlp.", layer_norm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])))
    return x


def diff_info_sharing(sd: SD, p: str, feats: List[Tensor], depth: int, heads: int, base: Optional[float] = 100.0, indices=None,
                      norm_intermediate: bool = True):
    """`DifferentialMultiViewCrossAttentionTransformer.forward` (diff_cross_attention_transformer.py:175-259; IFR :390-507).
    heads = the TRANSFORMER's num_heads; its blocks are built with heads // 2 (:110-113).  base None: no positional encoding."""
    nv = len(feats)
    B, _, h, w = feats[0].shape
    toks = [f.permute(0, 2, 3, 1).reshape(B, h * w, f.shape[1]) for f in feats]
    pos = [patch_positions(B, h, w, f.device) if base is not None else None for f in feats]
    if (p + "proj_embed.weight") in sd:
        toks = [linear(t, sd[p + "proj_embed.weight"], sd[p + "proj_embed.bias"]) for t in toks]
    take = feature_take_indices(depth, indices)[0] if indices is not None else []
    inter = []
    nw, nb = sd[p + "norm.weight"], sd[p + "norm.bias"]
    for k in range(depth):
        new = []
        for v in range(nv):
            others = torch.cat([toks[i] for i in range(nv) if i != v], dim=1)
            opos = torch.cat([pos[i] for i in range(nv) if i != v], dim=1) if base is not None else None
            new.append(diff_decoder_block(sd, f"{p}multi_view_branches.{v}.{k}.", toks[v], others, pos[v], opos, heads // 2, k,
                                          base if base is not None else 100.0))
        toks = new
        if k in take:
            inter.append([layer_norm(t, nw, nb) if norm_intermediate else t for t in toks])
    dim = toks[0].shape[-1]

    def to_bchw(t):
        return t.reshape(B, h, w, dim).permute(0, 3, 1, 2).contiguous()

    out = [to_bchw(layer_norm(t, nw, nb)) for t in toks]
    if indices is not None:
        return out, [[to_bchw(t) for t in lvl] for lvl in inter]
    return out


def patch_embed(sd: SD, p: str, img: Tensor, patch: int, true_shape: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Conv2d(3,C,k=s=patch)+bias, flatten(2).transpose(1,2) (libs/croco/patch_embed.py:68-82),
    restated as unfold + matmul (kernel == stride so patches do not overlap).
    true_shape [B,2] (height, width): `ManyAR_PatchEmbed` (patch_embed.py:85-127) -- samples whose true width < height
    were stored transposed (landscape); their patches are taken from the image with the last two axes swapped and their
    positions come from the (W/p, H/p) grid."""
    B, Cin, H, W = img.shape
    assert H % patch == 0 and W % patch == 0
    h, w = H // patch, W // patch
    wgt = sd[p + "proj.weight"]
    wmat = wgt.reshape(wgt.shape[0], -1).t()

    def embed(im, hh, ww):
        if _FUNCTIONAL:  # nn.Conv2d(3, C, k = s = patch) then flatten(2).transpose(1, 2), patch_embed.py:77-79
            return F.conv2d(im, wgt, sd[p + "proj.bias"], stride=patch).flatten(2).transpose(1, 2)
        cols = im.reshape(im.shape[0], Cin, hh, patch, ww, patch).permute(0, 2, 4, 1, 3, 5).reshape(im.shape[0], hh * ww, Cin * patch * patch)
        return cols @ wmat + sd[p + "proj.bias"]

    if true_shape is None:
        return embed(img, h, w), patch_positions(B, h, w, img.device)
    assert W >= H, f"img should be in landscape mode, but got W={W} H={H}"
    assert tuple(true_shape.shape) == (B, 2)
    portrait = (true_shape[:, 1] < true_shape[:, 0]).tolist()
    xs, pos = [], []
    for b_, is_p in enumerate(portrait):
        if is_p:
            xs.append(embed(img[b_:b_ + 1].swapaxes(-1, -2), w, h))
            pos.append(patch_positions(1, w, h, img.device))
        else:
            xs.append(embed(img[b_:b_ + 1], h, w))
            pos.append(patch_positions(1, h, w, img.device))
    return torch.cat(xs, 0), torch.cat(pos, 0)


def croco_encoder(
    sd: SD, p: str, img: Tensor, depth: int, heads: int, patch: int = 16, base: float = 100.0,
    indices=None, norm_intermediate: bool = True, true_shape: Optional[Tensor] = None,
):
    """encoders/croco.py:147-182 (and the IFR variant :260-327 when `indices` is given).
    Returns BCHW features (and a list of intermediate BCHW features).  `true_shape`: ManyAR patch-embed (croco.py:160-168)."""
    B, _, H, W = img.shape
    x, pos = patch_embed(sd, p + "patch_embed.", img, patch, true_shape)
    take = feature_take_indices(depth, indices)[0] if indices is not None else []
    inter = []
    for i in range(depth):
        x = encoder_block(sd, f"{p}enc_blocks.{i}.", x, pos, heads, base)
        if i in take:
            inter.append(layer_norm(x, sd[p + "enc_norm.weight"], sd[p + "enc_norm.bias"]) if norm_intermediate else x)
    x = layer_norm(x, sd[p + "enc_norm.weight"], sd[p + "enc_norm.bias"])
    C = x.shape[-1]

    def to_bchw(t):
        return t.permute(0, 2, 1).reshape(B, C, H // patch, W // patch).contiguous()

    if indices is not None:
        return to_bchw(x), [to_bchw(t) for t in inter]
    return to_bchw(x)


def info_sharing(
    sd: SD, p: str, feats: List[Tensor], depth: int, heads: int, base: float = 100.0,
    indices=None, norm_intermediate: bool = True, softmax_scaling=None,
):
    """info_sharing/cross_attention_transformer.py:191-275 (IFR: :390-505).  Each view's block
    at depth k reads the *other* views' tokens from depth k-1."""
    nv = len(feats)
    B, _, h, w = feats[0].shape
    toks = [f.permute(0, 2, 3, 1).reshape(B, h * w, f.shape[1]) for f in feats]
    pos = [patch_positions(B, h, w, f.device) for f in feats]
    if (p + "proj_embed.weight") in sd:
        toks = [linear(t, sd[p + "proj_embed.weight"], sd[p + "proj_embed.bias"]) for t in toks]
    take = feature_take_indices(depth, indices)[0] if indices is not None else []
    inter = []
    nw, nb = sd[p + "norm.weight"], sd[p + "norm.bias"]
    for k in range(depth):
        new = []
        for v in range(nv):
            others = torch.cat([toks[i] for i in range(nv) if i != v], dim=1)
            opos = torch.cat([pos[i] for i in range(nv) if i != v], dim=1)
            new.append(decoder_block(sd, f"{p}multi_view_branches.{v}.{k}.", toks[v], others, pos[v], opos, heads, base,
                                     softmax_scaling))
        toks = new
        if k in take:
            inter.append([layer_norm(t, nw, nb) if norm_intermediate else t for t in toks])
    dim = toks[0].shape[-1]

    def to_bchw(t):
        return t.reshape(B, h, w, dim).permute(0, 3, 1, 2).contiguous()

    out = [to_bchw(layer_norm(t, nw, nb)) for t in toks]
    if indices is not None:
        return out, [[to_bchw(t) for t in lvl] for lvl in inter]
    return out


def linear_head(sd: SD, p: str, feat: Tensor, patch: int) -> Tensor:
    """1x1 conv C -> out*p^2 then pixel_shuffle(p) (prediction_heads/linear.py:47-54, :81-82)."""
    w = sd[p + "linear.weight"]
    y = torch.einsum("bchw,oc->bohw", feat, w.reshape(w.shape[0], -1)) + sd[p + "linear.bias"][None, :, None, None]
    return pixel_shuffle(y, patch)


def pointmap_conf_adaptor(x: Tensor, depth_mode=("exp", -math.inf, math.inf), conf_mode=("exp", 1.0, math.inf)):
    """PointMapWithConfidenceAdaptor: prediction_heads/adaptors.py:1217-1230 -> PointMapAdaptor
    :318-355 (exp: xyz/clip(d,1e-8)*expm1(d)) + ConfidenceAdaptor :1068-1083 (vmin + exp(x).clip(max=vmax-vmin))."""
    xyz, c = x[:, :3], x[:, 3:4]
    mode, vmin, vmax = depth_mode
    if mode == "linear":
        pts = xyz
    else:
        d = xyz.norm(dim=1, keepdim=True)
        unit = xyz / d.clip(min=1e-8)
        if mode == "exp":
            pts = unit * torch.expm1(d)
        elif mode == "square":
            pts = unit * d.square()
        else:
            raise ValueError(mode)
    if not (vmin == -math.inf and vmax == math.inf):
        pts = pts.clip(vmin, vmax)
    cmode, cmin, cmax = conf_mode
    assert cmode == "exp"
    conf = cmin + c.exp().clip(max=cmax - cmin)
    return pts, conf


def depth_adaptor(x: Tensor, mode: str = "exp", vmin=-math.inf, vmax=math.inf) -> Tensor:
    """prediction_heads/adaptors.py:233-257."""
    if mode == "exp":
        y = torch.exp(x)
    elif mode == "square":
        y = x * x
    else:
        y = x
    if not (vmin == -math.inf and vmax == math.inf):
        y = y.clip(vmin, vmax)
    return y


def view_sinusoid_table(n_position: int, d_hid: int, base: float = 10000.0) -> Tensor:
    """info_sharing/global_attention_transformer.py:198-208 (`_get_sinusoid_encoding_table`), computed in float64 like the
    reference's numpy code, returned as fp32."""
    import numpy as np

    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(base, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).float()


def self_attention_info_sharing(sd: SD, p: str, feats: List[Tensor], depth: int, heads: int, *, alternating: bool = False,
                                base: Optional[float] = None, distinguish_ref: bool = True, pe_for_non_ref: bool = True,
                                max_num_views_for_pe: int = 1000, softmax_scaling=None, indices: Optional[Sequence[int]] = None,
                                norm_intermediate: bool = True, extra: Optional[Tensor] = None,
                                extra_per_view: Optional[List[Tensor]] = None):
    """`MultiViewGlobalAttentionTransformer.forward` (global_attention_transformer.py:224-462) and, with
    `alternating=True`, `MultiViewAlternatingAttentionTransformer.forward` (alternating_attention_transformer.py:397-447:
    even depths attend over all tokens, odd depths inside each view), for `use_rand_idx_pe_for_non_reference_views=False`.
    Blocks are `SelfAttentionBlock`s (utils/transformer_blocks.py:415-514).  base: RoPE frequency base when
    `custom_positional_encoding="rope"`, else None.
    extra [B, C, T] / extra_per_view (V x [B, C, Tv]): additional input tokens (:266-333).  Per-view extras follow their
    view's patch tokens and receive its view encoding; global extras close the sequence, get no view encoding and skip the
    frame-level blocks of the alternating variant.
    indices: the IFR variants (global_attention_transformer.py:766-774, :880-897): also collect `norm(x)` (or x) after those
    depths.  Returns per-view maps -- or, with additional tokens, the triple (maps, global token features [B, dim, T] or None,
    per-view token features or None) -- and with `indices` the pair (that for the final output, [that per taken depth])."""
    V = len(feats)
    B, C_in, h, w = feats[0].shape
    N = h * w
    assert base is None or (extra is None and extra_per_view is None), "no positional encoding with additional tokens (:341-351)"
    seqs = []
    for v in range(V):
        t = feats[v].reshape(B, C_in, N)
        if extra_per_view is not None:
            t = torch.cat([t, extra_per_view[v]], dim=2)
        seqs.append(t.permute(0, 2, 1))
    Np = seqs[0].shape[1]
    T = extra.shape[2] if extra is not None else 0
    x = torch.cat(seqs + ([extra.permute(0, 2, 1)] if extra is not None else []), dim=1)
    if p + "proj_embed.weight" in sd:
        x = linear(x, sd[p + "proj_embed.weight"], sd[p + "proj_embed.bias"])
    dim = x.shape[-1]
    if distinguish_ref:
        tab = view_sinusoid_table(max_num_views_for_pe if pe_for_non_ref else 1, dim).to(x.device)
        pe = torch.zeros(V, dim, device=x.device)
        pe[0] = tab[0]
        if pe_for_non_ref:
            pe[1:] = tab[1:V]
        pe = torch.cat([pe.repeat_interleave(Np, dim=0), torch.zeros(T, dim, device=x.device)], dim=0)
        x = x + pe[None]
    pos = patch_positions(B, h, w, x.device).repeat(1, V, 1) if base is not None else None
    inter = []
    for i in range(depth):
        bp = f"{p}self_attention_blocks.{i}."
        if alternating and i % 2 == 1:
            xf = x[:, :V * Np].reshape(B * V, Np, dim)
            pf = pos.reshape(B * V, N, 2) if pos is not None else None
            xf = encoder_block(sd, bp, xf, pf, heads, base, softmax_scaling).reshape(B, V * Np, dim)
            x = torch.cat([xf, x[:, V * Np:]], dim=1)
        else:
            x = encoder_block(sd, bp, x, pos, heads, base, softmax_scaling)
        if indices is not None and i in indices:
            inter.append(layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"]) if norm_intermediate else x)
    x = layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"])

    def views(t):
        tv = t[:, :V * Np].reshape(B, V, Np, dim)
        maps = [tv[:, v, :N].reshape(B, h, w, dim).permute(0, 3, 1, 2).contiguous() for v in range(V)]
        if extra is None and extra_per_view is None:
            return maps
        pv = [tv[:, v, N:].permute(0, 2, 1).contiguous() for v in range(V)] if extra_per_view is not None else None
        ex = t[:, V * Np:].permute(0, 2, 1).contiguous() if extra is not None else None
        return maps, ex, pv

    if indices is None:
        return views(x)
    return views(x), [views(t) for t in inter]

# --------------------------------------------------------------------------------------
# DPT head (prediction_heads/dpt.py:94-232, :271-311; libs/croco/dpt_block.py:114-255)
# --------------------------------------------------------------------------------------
def _rcu(sd: SD, p: str, z: Tensor) -> Tensor:
    """ResidualConvUnit_custom: z + conv3x3(relu(conv3x3(relu(z)))) (dpt_block.py:114-177, ReLU not in-place)."""
    o = F.conv2d(F.relu(z), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    o = F.conv2d(F.relu(o), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return o + z


def _fusion(sd: SD, p: str, a: Tensor, b: Optional[Tensor]) -> Tensor:
    """FeatureFusionBlock_custom (dpt_block.py:225-255): a [+ RCU1(b)] -> RCU2 -> bilinear x2 (align_corners) -> 1x1."""
    out = a
    if b is not None:
        out = out + _rcu(sd, p + "resConfUnit1.", b)
    out = _rcu(sd, p + "resConfUnit2.", out)
    out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def dpt_feature(sd: SD, p: str, feats: List[Tensor]) -> Tensor:
    """DPTFeature.forward (dpt.py:180-232) for hooks [0,1,2,3]."""
    layers = []
    for j, f in enumerate(feats):
        q = f"{p}input_process.{j}.0."
        x = F.conv2d(f, sd[q + "0.weight"], sd[q + "0.bias"])  # 1x1 C_j -> L_j
        if j == 0:
            x = F.conv_transpose2d(x, sd[q + "1.weight"], sd[q + "1.bias"], stride=4)
        elif j == 1:
            x = F.conv_transpose2d(x, sd[q + "1.weight"], sd[q + "1.bias"], stride=2)
        elif j == 3:
            x = F.conv2d(x, sd[q + "1.weight"], sd[q + "1.bias"], stride=2, padding=1)
        layers.append(F.conv2d(x, sd[f"{p}scratch.layer_rn.{j}.weight"], None, padding=1))
    l0, l1, l2, l3 = layers
    p4 = _fusion(sd, p + "scratch.refinenet4.", l3, None)[:, :, : l2.shape[2], : l2.shape[3]]
    p3 = _fusion(sd, p + "scratch.refinenet3.", p4, l2)
    p2 = _fusion(sd, p + "scratch.refinenet2.", p3, l1)
    return _fusion(sd, p + "scratch.refinenet1.", p2, l0)


def dpt_regressor(sd: SD, p: str, x: Tensor, out_hw: Tuple[int, int]) -> Tensor:
    """DPTRegressionProcessor.forward (dpt.py:285-311)."""
    x = F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    x = F.interpolate(x, size=out_hw, mode="bilinear", align_corners=True)
    x = F.relu(F.conv2d(x, sd[p + "conv2.0.weight"], sd[p + "conv2.0.bias"], padding=1))
    return F.conv2d(x, sd[p + "conv2.2.weight"], sd[p + "conv2.2.bias"])


# --------------------------------------------------------------------------------------
# whole model: factory/dust3r.py:250-332
# --------------------------------------------------------------------------------------
def dust3r_forward(
    sd: SD, img1: Tensor, img2: Tensor, *, enc_depth=24, enc_heads=16, dec_depth=12, dec_heads=12,
    patch=16, base=100.0, head="linear", instances=None,
) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    B, _, H, W = img1.shape
    sym = instances is not None and is_symmetrized(instances[0], instances[1])
    a, b = (img1[::2], img2[::2]) if sym else (img1, img2)
    if head == "linear":
        feat = croco_encoder(sd, "encoder.", torch.cat((a, b), 0), enc_depth, enc_heads, patch, base)
    else:
        feat = croco_encoder(sd, "encoder.", torch.cat((a, b), 0), enc_depth, enc_heads, patch, base)
    f1, f2 = feat.chunk(2, dim=0)
    if sym:
        f1, f2 = interleave(f1, f2)
    # heads + adaptors: fp32 inputs, autocast disabled (factory/dust3r.py:285-309)
    if head == "linear":
        d1, d2 = info_sharing(sd, "info_sharing.", [f1, f2], dec_depth, dec_heads, base)
        with torch.autocast(img1.device.type, enabled=False):
            o1 = linear_head(sd, "head1.", d1.float(), patch)
            o2 = linear_head(sd, "head2.", d2.float(), patch)
    else:
        (d1, d2), inter = info_sharing(sd, "info_sharing.", [f1, f2], dec_depth, dec_heads, base,
                                       indices=[5, 8], norm_intermediate=False)
        with torch.autocast(img1.device.type, enabled=False):
            h1 = [t.float() for t in (f1, inter[0][0], inter[1][0], d1)]
            h2 = [t.float() for t in (f2, inter[0][1], inter[1][1], d2)]
            o1 = dpt_regressor(sd, "dpt_regressor_head1.", dpt_feature(sd, "dpt_feature_head1.", h1), (H, W))
            o2 = dpt_regressor(sd, "dpt_regressor_head2.", dpt_feature(sd, "dpt_feature_head2.", h2), (H, W))
    with torch.autocast(img1.device.type, enabled=False):
        p1, c1 = pointmap_conf_adaptor(o1)
        p2, c2 = pointmap_conf_adaptor(o2)
    res1 = {"pts3d": p1.permute(0, 2, 3, 1).contiguous(), "conf": c1.permute(0, 2, 3, 1).contiguous()}
    res2 = {"pts3d_in_other_view": p2.permute(0, 2, 3, 1).contiguous(), "conf": c2.permute(0, 2, 3, 1).contiguous()}
    return res1, res2


def bench_loss(res1, res2) -> Tensor:
    """The `.sum().backward()` idiom of encoders/utils.py:29-31 applied to all four outputs."""
    return res1["pts3d"].sum() + res1["conf"].sum() + res2["pts3d_in_other_view"].sum() + res2["conf"].sum()


def seeded_state_dict(shapes: Dict[str, Sequence[int]], seed: int, dtype=torch.float32) -> SD:
    """Deterministic, machine-independent weights for goldens: numpy RandomState stream in key order.
    Matrices ~ U(-a, a) with the xavier bound, biases small non-zero, LayerNorm weights near 1."""
    import numpy as np

    rs = np.random.RandomState(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if len(shp) >= 2:
            fan_out, fan_in = shp[0], int(np.prod(shp[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            v = rs.uniform(-a, a, size=shp)
        elif "norm" in k and k.endswith("weight"):
            v = 1.0 + 0.1 * rs.standard_normal(shp)
        elif k.endswith("gamma"):  # LayerScale: O(1) so that the scaled branch matters in the parity metric
            v = 1.0 + 0.3 * rs.standard_normal(shp)
        elif k.endswith("subln.weight"):  # RMS sub-layer norm of the Diff layers
            v = 1.0 + 0.1 * rs.standard_normal(shp)
        elif ".lambda_" in k:  # DiffAttention lambdas: large enough that exp(sum(lq * lk)) moves away from 1
            v = 0.3 * rs.standard_normal(shp)
        else:
            v = 0.02 * rs.standard_normal(shp)
        sd[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
    return sd
