"""GPU: fixed cost of one uc_gemm launch (pair kernel) -- tiny problems, back-to-back launches with PDL, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniception_b200 import ops


def t(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (m, n, k) in ((256, 256, 64), (256, 256, 768), (18944, 256, 64), (18944, 256, 768), (8192, 768, 64), (8192, 768, 768), (8192, 768, 3072),
                  (3136, 1024, 1024), (3136, 3072, 1024)):
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") / 30).bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(n, device="cuda")
    g = torch.cuda.CUDAGraph()
    fn = lambda: ops.gemm(x, w, out, bias=bias)
    fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    us_graph = t(g.replay, 20) / 20
    print(f"uc_gemm {m}x{n}x{k}: eager back-to-back {t(fn):6.1f} us / launch; inside a CUDA graph of 20 {us_graph:6.1f} us / launch "
          f"({2.0 * m * n * k / us_graph / 1e6:6.0f} TFLOP/s)")
empty = torch.empty(1, device="cuda")
print(f"torch elementwise add (1 element), eager: {t(lambda: empty.add_(1.0)):6.1f} us / launch")

if "--trace" in sys.argv:
    import ctypes as C
    from uniception_b200 import _lib
    fn = _lib.lib.uc_debug_set_gemm_trace
    fn.argtypes = [C.c_void_p]
    for (m, n, k) in ((18944, 256, 64), (8192, 768, 768), (16384, 1024, 1024)):
        x = torch.randn(m, k, device="cuda").bfloat16()
        w = (torch.randn(n, k, device="cuda") / 30).bfloat16()
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        bias = torch.zeros(n, device="cuda")
        for _ in range(5):
            ops.gemm(x, w, out, bias=bias)
        bufs = [torch.zeros(148 * 16, dtype=torch.int64, device="cuda") for _ in range(3)]
        torch.cuda.synchronize()
        for b_ in bufs:  # three consecutive launches, each with its own trace buffer (set between launches: host-ordered)
            assert fn(b_.data_ptr()) == 0
            ops.gemm(x, w, out, bias=bias)
        torch.cuda.synchronize()
        assert fn(None) == 0
        t = torch.stack([b_.cpu().view(148, 16) for b_ in bufs])  # [launch, cta, event]
        names = ["entry", "setup", "dep_wait", "operands", "acc_full", "epi_done", "exit"]
        for li in (1, 2):
            t0 = int(t[li, :, 0].min())
            prev_exit = int(t[li - 1, :, 6].max())
            print(f"uc_gemm {m}x{n}x{k} launch {li}: first CTA entry {t0 - prev_exit:+d} ns after the previous launch's last exit; "
                  f"last CTA entry +{int(t[li, :, 0].max()) - t0} ns")
            for cta in (0, 147):
                print("   cta %3d: " % cta + "  ".join(f"{nm} +{int(t[li, cta, e]) - t0}" for e, nm in enumerate(names)))
            print("   cta 147 epilogue of tile 0 (warp 2): " + "  ".join(f"c{c}: regs +{int(t[li, 147, 7 + 2 * c]) - t0} staged +{int(t[li, 147, 8 + 2 * c]) - t0}" for c in range(4))
                  + f"  loop left +{int(t[li, 147, 15]) - t0}  stores complete +{int(t[li, 147, 5]) - t0}")
            print(f"   all CTAs: exit max +{int(t[li, :, 6].max()) - t0} ns, acc_full median +{int(t[li, :, 4].median()) - t0}, epi_done median +{int(t[li, :, 5].median()) - t0}")
