"""One forward / dgrad / wgrad launch of uc_conv3x3 at the C5 refinenet shape (8 x 148 x 148, 256 -> 256): the target of
`ncu --set full -k regex:gemm2_kernel` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uniception_b200 import ops

B, H, W, ci, co = 8, 148, 148, 256, 256
x = torch.randn(B * H * W, ci, device="cuda").bfloat16()
w16 = (torch.randn(co, 9 * ci, device="cuda") / 50).bfloat16()
gy = torch.randn(B * H * W, co, device="cuda").bfloat16()
dw = torch.zeros(co, 9 * ci, device="cuda")
bias = torch.zeros(co, device="cuda")
for _ in range(2):
    ops.conv3x3_fwd(x, w16, B, H, W, bias=bias)
    ops.conv3x3_dgrad(gy, w16, B, H, W)
    ops.conv3x3_wgrad_(x, gy, dw, B, H, W)
torch.cuda.synchronize()
print("done")
