"""Manual measurement (not collected by pytest): the fp32 oracle restatement of the reference path run EAGERLY on the
GPU under `torch.autocast(bf16)` -- torch SDPA / cuBLAS / eager elementwise, i.e. what the reference's own PyTorch path
does on a GPU (the GPU-eager denominator of the "6x" target).  The reference tree itself does not exist on
the GPU box, so its restatement stands in; the product never imports this.

    python tests/measure_gpu_eager_oracle.py [--pairs 8] [--size 512] [--steps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch

import dust3r_oracle as O
import uniception_b200 as U

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
B, S = args.pairs, args.size
torch.manual_seed(42)
m = U.DUSt3R(name="dust3r", img_size=(S, S))  # parameter container only (CPU construction, reference init)
sd = {k: v.detach().cuda().requires_grad_(True) for k, v in m.state_dict().items()}
del m
g = torch.Generator().manual_seed(1234)
a = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1).cuda()
b = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1).cuda()


def step():
    for v in sd.values():
        v.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        r1, r2 = O.dust3r_forward(sd, a, b)
    O.bench_loss(r1, r2).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(json.dumps({"what": "oracle restatement, GPU eager, bf16 autocast (torch SDPA + cuBLAS)", "pairs_per_step": B, "size": S,
                  "ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "torch": torch.__version__,
                  "sdpa_backends": {"flash": torch.backends.cuda.flash_sdp_enabled(), "mem_efficient": torch.backends.cuda.mem_efficient_sdp_enabled(),
                                    "cudnn": torch.backends.cuda.cudnn_sdp_enabled()}}))
