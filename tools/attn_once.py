"""One forward and one backward attention launch per bench shape (encoder 16x16x1024, decoder 8x12x1024): the target of
`ncu -k regex:attn_ --metrics sm__pipe_tensor_cycles_active...` captures (BASELINE.json's secondary metric)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniception_b200 import ops

torch.manual_seed(0)
for (B, H, N) in [(16, 16, 1024), (8, 12, 1024)]:
    C = H * 64
    qkv = torch.randn(B * N, 3 * C, device="cuda").bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    for _ in range(2):  # second pass = warm caches / TMA descriptors
        o, lse = ops.attn_fwd(q, k, v, B, H, N, N, 0.125)
        d_o = torch.randn_like(o)
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(q, k, v, o, d_o, lse, B, H, N, N, 0.125, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:])
    torch.cuda.synchronize()
print("done")
