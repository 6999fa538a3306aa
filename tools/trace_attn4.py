"""GPU bring-up aid: cycle trace of CTA 0 of attn_fwd4_kernel at the encoder shape (A3_TRACE events, clock64 deltas)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from uniception_b200 import _lib, ops

B, H, N = 16, 16, 1024
Cc = H * 64
qkv = torch.randn(B * N, 3 * Cc, device="cuda").bfloat16()
q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
for _ in range(3):
    ops.attn_fwd(q, k, v, B, H, N, N, 0.125)
buf = torch.zeros(3 * 64 * 16 + 4 * 296, dtype=torch.int64, device="cuda")
fn = _lib.lib.uc_debug_set_attn2_trace
fn.argtypes = [C.c_void_p]
assert fn(buf.data_ptr()) == 0
ops.attn_fwd(q, k, v, B, H, N, N, 0.125)
torch.cuda.synchronize()
assert fn(None) == 0
t = buf.cpu()[:3 * 64 * 16].view(3, 64, 16)
t0 = int(t[1, 0, 0])
print("tile | MMA: top sfreeA qkA sfreeB qkB pvA pvB | GROUP A half0: top S_ready loaded max bar exp0 o_done exp1+stored arrived [item end] | GROUP B half0: same")
for j in range(26):
    m = [int(x) - t0 for x in t[0, j, :7]]
    s0 = [int(x) - t0 for x in t[1, j, :10]]
    s1 = [int(x) - t0 for x in t[2, j, :10]]
    print(f"{j:3d} | " + " ".join(f"{x:6d}" for x in m) + " | " + " ".join(f"{x:6d}" for x in s0) + " | " + " ".join(f"{x:6d}" for x in s1))
print("per-tile period (group A S_ready to S_ready):", [int(t[1, j + 1, 1] - t[1, j, 1]) for j in range(8, 22)])
