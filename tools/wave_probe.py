import sys, torch
sys.path.insert(0, '/root/repo')
from uniception_b200 import ops
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (n, k) in ((1024, 4096), (1024, 1024), (768, 768), (768, 3072)):
    w = (torch.randn(n, k, device='cuda') / 30).bfloat16()
    bias = torch.zeros(n, device='cuda')
    for tiles_m in (37, 55.5, 64, 74, 92.5) if n == 1024 else (24 + 2/3, 32, 37, 49 + 1/3):
        m = int(round(tiles_m * 256))
        x = torch.randn(m, k, device='cuda').bfloat16()
        out = torch.empty(m, n, device='cuda', dtype=torch.bfloat16)
        us = t(lambda: ops.gemm(x, w, out, bias=bias))
        tiles = (m + 255) // 256 * (n // 256)
        print(f"n={n} k={k} m={m}: tiles {tiles} = {tiles/74:.2f} waves  {us:7.1f} us  {2.0*m*n*k/us/1e6:6.0f} TFLOP/s  per-wave {us/ -(-tiles//74):6.1f} us")
