"""A few launches of the hot GEMM shapes with their real epilogues (for ncu):
    ncu --set full --import-source on -k regex:gemm2 -c 6 -o gpurun_out/gemm python tools/gemm_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniception_b200 import ops

torch.manual_seed(0)
m, C, Hd = 16384, 1024, 4096
x = torch.randn(m, C, device="cuda").bfloat16()
w1 = (torch.randn(Hd, C, device="cuda") * 0.03).bfloat16()
b1 = torch.zeros(Hd, device="cuda")
act = torch.empty(m, Hd, device="cuda", dtype=torch.bfloat16)
pre = torch.empty_like(act)
dy = torch.randn(m, C, device="cuda").bfloat16()
w2 = (torch.randn(C, Hd, device="cuda") * 0.03).bfloat16()
dpre = torch.empty_like(act)
res = torch.randn(m, C, device="cuda").bfloat16()
out = torch.empty(m, C, device="cuda", dtype=torch.bfloat16)
b2 = torch.zeros(C, device="cuda")
for _ in range(2):
    ops.gemm(x, w1, act, bias=b1, gelu=True, aux_out=pre)          # fc1 + bias + GELU (epi 5)
    ops.gemm(dy, w2, dpre, b_layout=1, gelu_bwd=True, aux_in=pre)  # fc2 dgrad * GELU' (epi 8)
    ops.gemm(act, w2, out, bias=b2, residual=res)                  # fc2 + bias + residual (epi 17)
torch.cuda.synchronize()
print("done")
