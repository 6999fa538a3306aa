"""GPU: attention kernel timings at the bench shapes, next to torch SDPA forward AND backward (same box, same process).

    python tools/attn_bench.py            # current kernels
    UC_ATTN_FWD=1 UC_ATTN_BWD=1 python tools/attn_bench.py   # first-generation kernels (A/B)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from uniception_b200 import ops


def _time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    tag = f"fwd={os.environ.get('UC_ATTN_FWD', '2')} bwd={os.environ.get('UC_ATTN_BWD', '2')} poly={os.environ.get('UC_ATTN_POLY', '1')}"
    for (B, H, N) in [(16, 16, 1024), (8, 12, 1024), (16, 16, 196), (8, 16, 1369)]:
        Cc = H * 64
        torch.manual_seed(0)
        qkv = torch.randn(B * N, 3 * Cc, device="cuda").bfloat16()
        q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
        o, lse = ops.attn_fwd(q, k, v, B, H, N, N, 0.125)
        fl = 4.0 * B * H * N * N * 64
        ms = _time(lambda: ops.attn_fwd(q, k, v, B, H, N, N, 0.125, out=o))
        print(f"[attn {tag}] fwd B{B} H{H} N{N}: {ms*1e3:.1f} us = {fl/ms/1e9:.0f} TFLOP/s", flush=True)
        do = torch.randn_like(o)
        dqkv = torch.empty_like(qkv)
        ms = _time(lambda: ops.attn_bwd(q, k, v, o, do, lse, B, H, N, N, 0.125, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:]))
        print(f"[attn {tag}] bwd (all kernels) B{B} H{H} N{N}: {ms*1e3:.1f} us = {2.5*fl/ms/1e9:.0f} TFLOP/s (5-GEMM convention)", flush=True)
        if os.environ.get("UC_ATTN_TORCH", "1") == "1":
            qh = q.reshape(B, N, H, 64).transpose(1, 2).detach().requires_grad_(True)
            kh = k.reshape(B, N, H, 64).transpose(1, 2).detach().requires_grad_(True)
            vh = v.reshape(B, N, H, 64).transpose(1, 2).detach().requires_grad_(True)
            ms_t = _time(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
            out = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
            g = torch.randn_like(out)
            ms_b = _time(lambda: torch.autograd.grad(out, (qh, kh, vh), g, retain_graph=True))
            print(f"[attn torch SDPA] fwd {ms_t*1e3:.1f} us = {fl/ms_t/1e9:.0f} TFLOP/s; bwd {ms_b*1e3:.1f} us = {2.5*fl/ms_b/1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
