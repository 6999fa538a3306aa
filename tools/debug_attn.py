"""debug helper: run one attention bwd case per subprocess with a short timeout"""
import math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))

def one(B, H, Nq, Nk):
    import torch
    from uniception_b200 import ops
    torch.manual_seed(6)
    Cc = H * 64
    q = torch.randn(B * Nq, Cc, device="cuda").bfloat16()
    k = torch.randn(B * Nk, Cc, device="cuda").bfloat16()
    v = torch.randn(B * Nk, Cc, device="cuda").bfloat16()
    do = torch.randn(B * Nq, Cc, device="cuda").bfloat16()
    o, lse = ops.attn_fwd(q, k, v, B, H, Nq, Nk, 0.125)
    torch.cuda.synchronize(); print("fwd done", flush=True)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attn_bwd(q, k, v, o, do, lse, B, H, Nq, Nk, 0.125, dq, dk, dv)
    torch.cuda.synchronize(); print("bwd done", flush=True)
    qf = q.float().reshape(B, Nq, H, 64).transpose(1, 2).requires_grad_(True)
    kf = k.float().reshape(B, Nk, H, 64).transpose(1, 2).requires_grad_(True)
    vf = v.float().reshape(B, Nk, H, 64).transpose(1, 2).requires_grad_(True)
    s = (qf @ kf.transpose(-2, -1)) * 0.125
    (s.softmax(-1) @ vf).backward(do.float().reshape(B, Nq, H, 64).transpose(1, 2))
    for n, a, g in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        ref = g.transpose(1, 2).reshape(a.shape)
        print(n, "rel", float((a.float() - ref).norm() / ref.norm()), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*map(int, sys.argv[1:5]))
    else:
        for case in [(1, 1, 128, 128), (1, 1, 64, 64), (2, 2, 256, 256), (1, 2, 100, 300), (2, 4, 1024, 1024)]:
            print("== case", case, flush=True)
            try:
                subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(c) for c in case], timeout=40)
            except subprocess.TimeoutExpired:
                print("   TIMEOUT", flush=True)
