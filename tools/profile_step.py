"""One profiled fwd+bwd step of the bench workload for ncu (use with --profile-from-start off):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--pairs 8] [--size 512]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import uniception_b200 as U

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--workload", default="linear", choices=["linear", "dpt", "c5"], help="linear / dpt: DUSt3R with that head; c5: ViT-L/14 + DPT depth")
args = ap.parse_args()
S, B = args.size, args.pairs
torch.manual_seed(42)
if args.workload == "c5":
    m = U.ViTDPTDepth(img_size=(S, S)).cuda()
else:
    m = U.DUSt3R(name="dust3r", img_size=(S, S), pred_head_type=args.workload).cuda()
pk = m.pack()
g = torch.Generator().manual_seed(1234)
a = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1).cuda()
b = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1).cuda()


def step():
    pk.zero_grad()
    if args.workload == "c5":
        m(a).sum().backward()
        return
    r1, r2 = m({"img": a, "instance": [str(i) for i in range(B)], "data_norm_type": "dust3r"},
               {"img": b, "instance": [str(B + i) for i in range(B)], "data_norm_type": "dust3r"})
    (r1["pts3d"].sum() + r1["conf"].sum() + r2["pts3d_in_other_view"].sum() + r2["conf"].sum()).backward()


for _ in range(args.warm):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
