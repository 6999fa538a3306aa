"""GPU bring-up aid: cycle trace of CTA (0,0) of attn_fwd2_kernel at the encoder shape (A2_TRACE events, clock64 deltas)."""
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
c = buf.cpu()[3 * 64 * 16:].view(296, 4)
t = buf.cpu()[:3 * 64 * 16].view(3, 64, 16)
t0 = int(t[1, 0, 0])
print("tile | MMA: top s_free_seen qk_issued p_ready_seen pv_issued | SOFTMAX half0: top s_full max_done bar_done exp0 o_done_seen s_free_arrived exp1 p_ready_arrived | half1: same")
for j in range(24):
    m = [int(x) - t0 for x in t[0, j, :5]]
    s0 = [int(x) - t0 for x in t[1, j, :9]]
    s1 = [int(x) - t0 for x in t[2, j, :9]]
    print(f"{j:3d} | " + " ".join(f"{x:6d}" for x in m) + " | " + " ".join(f"{x:6d}" for x in s0) + " | " + " ".join(f"{x:6d}" for x in s1))

for j in (7, 15, 23):
    print(f"epilogue after tile {j}: half0", [int(x) - t0 for x in t[1, j, 8:15]], " half1", [int(x) - t0 for x in t[2, j, 8:15]],
          "(p_ready_arrived, top, o_done_seen, bar, O stored, lse stored, bar)")
st = c[:, 0] - c[:, 0].min()
en = c[:, 1] - c[:, 0].min()
dur = (c[:, 1] - c[:, 0])
print("per-CTA (softmax warp 2) ns: start min/max", int(st.min()), int(st.max()), " end min/max", int(en.min()), int(en.max()),
      " duration min/median/max", int(dur.min()), int(dur.median()), int(dur.max()))
import collections
cnt = collections.Counter(int(x) for x in c[:, 2])
print("CTAs per SM histogram:", collections.Counter(cnt.values()), " distinct SMs:", len(cnt))
slow = sorted(range(296), key=lambda i: -int(dur[i]))[:8]
print("slowest CTAs (id, sm, dur ns):", [(i, int(c[i, 2]), int(dur[i])) for i in slow])
fast = sorted(range(296), key=lambda i: int(dur[i]))[:8]
print("fastest CTAs (id, sm, dur ns):", [(i, int(c[i, 2]), int(dur[i])) for i in fast])
