"""Cycle trace of one CTA of the attention-backward kernels (debug aid)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniception_b200 import _lib, ops
B, H, N = 16, 16, 1024
Cc = H * 64
qkv = torch.randn(B * N, 3 * Cc, device="cuda").bfloat16()
q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
o, lse = ops.attn_fwd(q, k, v, B, H, N, N, 0.125)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.attn_bwd(q, k, v, o, do, lse, B, H, N, N, 0.125, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:])
torch.cuda.synchronize()
trace = torch.zeros(2048, dtype=torch.int64, device="cuda")
fn = _lib.lib.uc_debug_set_trace
fn.argtypes = [ctypes.c_void_p]
fn(trace.data_ptr())
ops.attn_bwd(q, k, v, o, do, lse, B, H, N, N, 0.125, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:])
torch.cuda.synchronize()
fn(None)
tt = trace.cpu().view(2, 2, 64, 8)
for mode in (0, 1):
    t = tt[mode]
    base = int(t[0, 0, 0])
    print(f"MODE {mode}: tile | MMA: sfull_ok sdp_issued ds_ready_ok acc_issued | EXP: sdp_ok ld_done math_done st_done   (cycles since first event)")
    for i in range(4, 12):
        m = [int(x) - base for x in t[0, i, :4]]
        e = [int(x) - base for x in t[1, i, :4]]
        print(f"{i:3d} | {m[0]:7d} {m[1]:7d} {m[2]:7d} {m[3]:7d} | {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[3]:7d}")
