"""GPU: uc_conv3x3 (implicit-GEMM 3x3 convolution) forward / dgrad / wgrad vs torch fp32 conv2d on bf16-rounded operands,
and timing at the DPT head's real map sizes next to the materialised-column path (uc_im2col3x3 -> uc_gemm -> uc_col2im3x3).

    python tools/conv_probe.py [--time]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from uniception_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def nhwc(t):  # [B,C,H,W] -> [B*H*W, C] bf16
    B, C, H, W = t.shape
    return t.permute(0, 2, 3, 1).reshape(B * H * W, C).bfloat16().contiguous()


def nchw(t, B, H, W):
    return t.float().view(B, H, W, -1).permute(0, 3, 1, 2)


def check(B, H, W, ci, co, relu=False, residual=False, tol=6e-3):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H * 10 + W + ci + co)
    x = torch.randn(B, ci, H, W, device="cuda", generator=g).bfloat16().float()
    w = (torch.randn(co, ci, 3, 3, device="cuda", generator=g) / (3 * ci ** 0.5)).bfloat16().float()
    b = torch.randn(co, device="cuda", generator=g)
    res = torch.randn(B, co, H, W, device="cuda", generator=g).bfloat16().float() if residual else None
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, b, padding=1)
    if relu:
        ref = ref.relu()
    if residual:
        ref = ref + res
    w16 = w.permute(0, 2, 3, 1).reshape(co, 9 * ci).bfloat16().contiguous()
    y = ops.conv3x3_fwd(nhwc(x), w16, B, H, W, bias=b, relu=relu, residual=nhwc(res) if residual else None)
    e_f = rel(nchw(y, B, H, W), ref)
    gy = torch.randn(B, co, H, W, device="cuda", generator=g).bfloat16().float()
    if relu:
        gy = gy * (ref > 0)
    ref.backward(gy)
    gyt = nhwc(gy)
    dx = ops.conv3x3_dgrad(gyt, w16, B, H, W)
    e_dx = rel(nchw(dx, B, H, W), xr.grad)
    dw = torch.zeros(co, 9 * ci, device="cuda")
    ops.conv3x3_wgrad_(nhwc(x), gyt, dw, B, H, W)
    e_dw = rel(dw.view(co, 3, 3, ci).permute(0, 3, 1, 2), wr.grad)
    # dgrad with the fused ReLU mask of the producer
    mask_src = torch.randn(B, ci, H, W, device="cuda", generator=g).relu().bfloat16()
    dxm = ops.conv3x3_dgrad(gyt, w16, B, H, W, relu_out=nhwc(mask_src.float()))
    e_dxm = rel(nchw(dxm, B, H, W), xr.grad * (mask_src > 0))
    ok = max(e_f, e_dx, e_dxm) <= tol and e_dw <= 1e-3
    print(f"[{'OK ' if ok else 'BAD'}] conv3x3 B{B} {H}x{W} {ci}->{co} relu={int(relu)} res={int(residual)}: fwd {e_f:.2e} dx {e_dx:.2e} "
          f"dx*mask {e_dxm:.2e} dw {e_dw:.2e}")
    return ok


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def bench(B, H, W, ci, co):
    x = torch.randn(B * H * W, ci, device="cuda").bfloat16()
    w16 = (torch.randn(co, 9 * ci, device="cuda") / 50).bfloat16()
    gy = torch.randn(B * H * W, co, device="cuda").bfloat16()
    dw = torch.zeros(co, 9 * ci, device="cuda")
    bias = torch.zeros(co, device="cuda")
    fl = 2.0 * B * H * W * 9 * ci * co
    t_f = timeit(lambda: ops.conv3x3_fwd(x, w16, B, H, W, bias=bias))
    t_d = timeit(lambda: ops.conv3x3_dgrad(gy, w16, B, H, W))
    t_w = timeit(lambda: ops.conv3x3_wgrad_(x, gy, dw, B, H, W))

    def old_fwd():
        cols = ops.im2col3x3(x, B, H, W, 1)
        out = torch.empty(B * H * W, co, dtype=torch.bfloat16, device="cuda")
        ops.gemm(cols, w16, out, bias=bias)

    def old_bwd():
        cols = ops.im2col3x3(x, B, H, W, 1)
        ops.gemm(gy, cols, dw, a_layout=1, b_layout=1, atomic=True)
        dcols = torch.empty(B * H * W, 9 * ci, dtype=torch.bfloat16, device="cuda")
        ops.gemm(gy, w16, dcols, b_layout=1)
        ops.col2im3x3(dcols, B, H, W, 1)

    t_of, t_ob = timeit(old_fwd, 5), timeit(old_bwd, 5)
    print(f"conv3x3 B{B} {H}x{W} {ci}->{co}: implicit fwd {t_f:8.1f} us ({fl / t_f / 1e6:6.0f} TFLOP/s)  dgrad {t_d:8.1f} us ({fl / t_d / 1e6:6.0f})  "
          f"wgrad {t_w:8.1f} us ({fl / t_w / 1e6:6.0f})  | im2col path fwd {t_of:8.1f} us  bwd {t_ob:8.1f} us  "
          f"-> x{(t_of + t_ob) / (t_f + t_d + t_w):.2f}")


if __name__ == "__main__":
    ok = True
    ok &= check(2, 12, 10, 64, 128)
    ok &= check(1, 37, 37, 256, 256, relu=True)
    ok &= check(2, 74, 74, 256, 256, residual=True)
    ok &= check(1, 64, 64, 192, 256)
    ok &= check(1, 33, 40, 128, 128)
    ok &= check(1, 16, 16, 768, 256)
    ok &= check(1, 148, 148, 256, 128, relu=True)
    ok &= check(1, 130, 518, 128, 128)
    ok &= check(3, 5, 7, 384, 256)
    print("ALL OK" if ok else "FAILURES")
    if "--time" in sys.argv:
        for shp in ((8, 148, 148, 256, 256), (8, 296, 296, 256, 128), (8, 518, 518, 128, 128), (16, 128, 128, 256, 256),
                    (16, 256, 256, 256, 128), (16, 512, 512, 128, 128), (16, 32, 32, 768, 256)):
            bench(*shp)
    sys.exit(0 if ok else 1)
