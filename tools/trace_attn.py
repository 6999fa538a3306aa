"""Cycle trace of CTA (0,0) of the pipelined attention-backward kernel (bring-up aid).
    python tools/trace_attn.py            # encoder shape B16 H16 N1024
Columns are cycles since the CTA's first recorded event."""
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
trace = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
fn = _lib.lib.uc_debug_set_attn_trace
fn.argtypes = [ctypes.c_void_p]
fn(trace.data_ptr())
ops.attn_bwd(q, k, v, o, do, lse, B, H, N, N, 0.125, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:])
torch.cuda.synchronize()
fn(None)
t = trace.cpu().view(3, 64, 8)
base = int(t[t > 0].min())
print("sub | MMA: loop_top sdp(j+1)_issued ds_ready(j)_seen acc_issued | EXP(warp2): top sdp_full_seen ld_done math_done arrived | DRAIN(tile=j/2): top dq_full_seen stored")
for j in range(16):
    m = [int(x) - base for x in t[0, j, :4]]
    e = [int(x) - base for x in t[1, j, :5]]
    d = [int(x) - base for x in t[2, j // 2, :3]] if j % 2 == 1 else []
    print(f"{j:3d} | " + " ".join(f"{x:7d}" for x in m) + " | " + " ".join(f"{x:7d}" for x in e) + " | " + " ".join(f"{x:7d}" for x in d))
