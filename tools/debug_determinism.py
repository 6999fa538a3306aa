"""Run each backward-path op twice on identical inputs and report run-to-run differences (fraction of differing elements, rel-L2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from uniception_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def rep(name, a, b):
    a, b = a.float(), b.float()
    nd = float((a != b).float().mean())
    rel = float((a - b).norm() / b.norm().clamp_min(1e-30))
    print(f"{name:34s} differing {nd:.3e}  rel-L2 {rel:.3e}", flush=True)


for (B, H, N) in ((2, 4, 1728), (6, 4, 576), (8, 12, 1024)):
    C = H * 64
    qkv = torch.randn(B * N, 3 * C, device=dev).bfloat16()
    o, lse = ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, 0.125)
    d_o = torch.randn(B * N, C, device=dev).bfloat16()
    outs = []
    for _ in range(3):
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, d_o, lse, B, H, N, N, 0.125, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:])
        torch.cuda.synchronize()
        outs.append(dqkv)
    for nm, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        rep(f"attn_bwd B{B} H{H} N{N} {nm} run0-1", outs[0][:, sl], outs[1][:, sl])
        rep(f"attn_bwd B{B} H{H} N{N} {nm} run1-2", outs[1][:, sl], outs[2][:, sl])
    o2, lse2 = ops.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, 0.125)
    rep(f"attn_fwd B{B} H{H} N{N} o", o, o2)

rows, Cc = 3456, 256
x = torch.randn(rows, Cc, device=dev).bfloat16(); dy = torch.randn(rows, Cc, device=dev).bfloat16(); dres = torch.randn(rows, Cc, device=dev).bfloat16()
g = torch.randn(Cc, device=dev)
y, mean, rstd = ops.layernorm_fwd(x, g, g, 1e-6, torch.bfloat16)
r = []
for _ in range(2):
    dg, db, cs = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    r.append((ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db, dres=dres, dx_colsum=cs), dg, cs))
rep("layernorm_bwd dx", r[0][0], r[1][0]); rep("layernorm_bwd dgamma", r[0][1], r[1][1]); rep("layernorm_bwd colsum", r[0][2], r[1][2])
w = torch.randn(1024, 256, device=dev).bfloat16(); dyy = torch.randn(rows, 1024, device=dev).bfloat16(); pre = torch.randn(rows, 256, device=dev).bfloat16()
r = []
for _ in range(2):
    dx = torch.empty(rows, 256, device=dev, dtype=torch.bfloat16); cs = torch.zeros(256, device=dev)
    ops.gemm(dyy, w, dx, b_layout=1, gelu_bwd=True, aux_in=pre, c_colsum=cs)
    r.append((dx, cs))
rep("gemm dgrad+gelu' dx", r[0][0], r[1][0]); rep("gemm dgrad c_colsum", r[0][1], r[1][1])
