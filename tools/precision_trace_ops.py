"""GPU: per-OP error trace inside encoder blocks (finer than tools/precision_trace.py).

For chosen blocks of a ViT encoder, every tensor the engine keeps for the backward pass (LN1 output, packed post-RoPE q/k/v,
attention output, residual stream after the attention, LN2 output, fc1 pre-activation, GELU output, residual stream after
the MLP) is compared with the same tensor of the fp32 oracle, next to the error of the reference arithmetic under
torch.autocast(bf16).  Two variants of "ours": (a) the real forward (errors accumulate from block 0), (b) every block fed the
fp32 oracle's own block input rounded to bf16 (isolates what ONE block adds).

    python tools/precision_trace_ops.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch

import dust3r_oracle as O
import uniception_b200 as U
from uniception_b200 import engine as E

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda"


def oracle_block(sd, p, x, pos, heads, base=100.0):
    """encoder_block of the oracle with every intermediate returned (same calls, same order)."""
    out = {}
    B, N, C = x.shape
    h1 = O.layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    out["ln1"] = h1
    qkv = O.linear(h1, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, heads, C // heads).transpose(1, 3)
    q, k, v = [qkv[:, :, i] for i in range(3)]
    q, k = O.rope2d(q, pos, base), O.rope2d(k, pos, base)
    out["q"], out["k"], out["v"] = [t.transpose(1, 2).reshape(B, N, C) for t in (q, k, v)]
    o = O.sdpa(q, k, v).transpose(1, 2).reshape(B, N, C)
    out["attn_out"] = o
    x = x + O.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    out["x_after_attn"] = x
    h2 = O.layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    out["ln2"] = h2
    pre = O.linear(h2, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    out["fc1_pre"] = pre
    act = O.gelu_erf(pre)
    out["gelu"] = act
    x = x + O.linear(act, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    out["x_after_mlp"] = x
    return x, out


def ours_block(saved_block, x_next, C):
    (x, mean, rstd, h1, qkv, o, lse, scale, qk, z), (x1, mean2, rstd2, h2, pre, act, z2) = saved_block
    return {"ln1": h1, "q": qkv[:, :C], "k": qkv[:, C:2 * C], "v": qkv[:, 2 * C:], "attn_out": o, "x_after_attn": x1, "ln2": h2,
            "fc1_pre": pre, "gelu": act, "x_after_mlp": x_next}


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def trace(C, depth, heads, S, blocks, seed=42):
    torch.manual_seed(seed)
    enc = U.CroCoEncoder(name="e", data_norm_type="dust3r", img_size=(S, S), enc_embed_dim=C, enc_depth=depth, enc_num_heads=heads).to(DEV)
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(1, 3, S, S, generator=g).clamp_(-1, 1).to(DEV)
    sd = {"encoder." + k: v.detach() for k, v in enc.state_dict().items()}
    pk = enc._pack()
    with torch.no_grad():
        y, _, saved = E.encoder_fwd(pk, "", img, depth, heads, 16, 100.0)
        xs_ours = [saved["blocks"][i][0][0] for i in range(depth)] + [saved["final"][0]]  # block inputs, then the final-norm input
        with O.reference_functionals():
            x32, pos = O.patch_embed(sd, "encoder.patch_embed.", img, 16)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                x16, _ = O.patch_embed(sd, "encoder.patch_embed.", img, 16)
            print(f"--- ViT C={C} depth={depth} {S}x{S}; patch-embed: ours {rel(xs_ours[0], x32):.3e}  autocast {rel(x16, x32):.3e} "
                  f"(autocast stream dtype {x16.dtype})")
            for i in range(depth):
                p = f"encoder.enc_blocks.{i}."
                x32n, o32 = oracle_block(sd, p, x32, pos, heads)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    x16n, o16 = oracle_block(sd, p, x16, pos, heads)
                    # one block in isolation: the fp32 block input rounded to bf16 through the autocast arithmetic
                    _, o16_iso = oracle_block(sd, p, x32.to(torch.bfloat16), pos, heads)
                if i in blocks:
                    ours = ours_block(saved["blocks"][i], xs_ours[i + 1], C)
                    # ours in isolation: run this block of the engine on the fp32 oracle's block input rounded to bf16
                    bs = []
                    N = x32.shape[1]
                    rope = saved["rope"]
                    xi = x32.reshape(-1, C).to(torch.bfloat16).contiguous()
                    xa = E.self_attn_fwd(pk, f"enc_blocks.{i}.", xi, 1, N, heads, rope, "norm1", bs)
                    xm = E.mlp_fwd(pk, f"enc_blocks.{i}.", xa, "norm2", bs)
                    ours_iso = ours_block(bs, xm, C)
                    print(f"  block {i}: tensor          ours(accum)  autocast(accum) | ours(isolated) autocast(isolated)  [dtype autocast]")
                    for k in o32:
                        r = o32[k].reshape(-1, C if k not in ("fc1_pre", "gelu") else 4 * C)
                        print(f"    {k:14s} {rel(ours[k], r):.3e}    {rel(o16[k], r):.3e}     |  {rel(ours_iso[k], r):.3e}      {rel(o16_iso[k], r):.3e}"
                              f"    {str(o16[k].dtype).replace('torch.', '')}")
                x32, x16 = x32n, x16n


if __name__ == "__main__":
    trace(1024, 24, 16, 512, blocks=(0, 1, 11, 23))
