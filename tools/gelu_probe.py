"""GPU: timing of the GELU / GELU' epilogue GEMMs (fc1 forward, fc2 dgrad) at the encoder and decoder shapes, CUDA events inside a
CUDA graph of 10 launches.  (Used for the A/B of a sixteen-epilogue-warp instance of these two epilogues, which lost:
profiles/r02b_attention_design_log.txt, last section.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniception_b200 import ops


def t(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 100 * 1e3


tag = "gelu"
for (m, C, Hd) in ((16384, 1024, 4096), (8192, 768, 3072)):
    x = torch.randn(m, C, device="cuda").bfloat16()
    w1 = (torch.randn(Hd, C, device="cuda") * 0.03).bfloat16()
    b1 = torch.zeros(Hd, device="cuda")
    act, pre = torch.empty(m, Hd, device="cuda", dtype=torch.bfloat16), torch.empty(m, Hd, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(m, C, device="cuda").bfloat16()
    w2 = (torch.randn(C, Hd, device="cuda") * 0.03).bfloat16()
    dpre = torch.empty_like(act)
    fl = 2.0 * m * C * Hd
    us = t(lambda: ops.gemm(x, w1, act, bias=b1, gelu=True, aux_out=pre))
    print(f"[{tag}] fc1 + GELU   {m}x{Hd}x{C}: {us:7.1f} us = {fl / us / 1e6:5.0f} TFLOP/s")
    us = t(lambda: ops.gemm(dy, w2, dpre, b_layout=1, gelu_bwd=True, aux_in=pre))
    print(f"[{tag}] fc2 dgrad GELU' {m}x{Hd}x{C}: {us:7.1f} us = {fl / us / 1e6:5.0f} TFLOP/s")
