"""Stage-by-stage comparison of the DPT engine (dpt_engine.dpt_forward) against the fp32 oracle on the tiny golden
model (bring-up aid).  python tools/debug_dpt.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
import dust3r_oracle as O
import uniception_b200 as U
from uniception_b200 import dpt_engine as D, ops
from golden_utils import load, weights

DEV = "cuda"
cfg, a = load("dust3r_tiny_dpt")
print({k: v for k, v in cfg.items() if k not in ("shapes", "grad_keys")})
m = U.DUSt3R(name="t", img_size=tuple(cfg["hw"]), pred_head_type="dpt", pred_head_feature_dim=32,
             encoder_kwargs=dict(enc_embed_dim=cfg["C_enc"], enc_depth=cfg["enc_depth"], enc_num_heads=cfg["enc_heads"]),
             info_sharing_kwargs=dict(depth=cfg["dec_depth"], dim=cfg["C_dec"], num_heads=cfg["dec_heads"]),
             dpt_kwargs=dict(layer_dims=[12, 24, 48, 96]), dpt_indices=tuple(cfg["ifr_indices"]))
m.load_state_dict(weights(cfg))
m = m.to(DEV)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}  # aliases resolved as load_state_dict does
img1, img2 = a["img1"].to(DEV), a["img2"].to(DEV)
B, _, H, W = img1.shape
h, w = H // 16, W // 16
feat = O.croco_encoder(sd, "encoder.", torch.cat((img1, img2), 0), cfg["enc_depth"], cfg["enc_heads"])
f1, f2 = feat.chunk(2, dim=0)
(d1, d2), inter = O.info_sharing(sd, "info_sharing.", [f1, f2], cfg["dec_depth"], cfg["dec_heads"], indices=cfg["ifr_indices"], norm_intermediate=False)
feats = [f1, inter[0][0], inter[1][0], d1]
print("oracle hook shapes", [tuple(f.shape) for f in feats])


def nlc(x):  # NCHW fp32 -> token-major bf16
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).bfloat16().contiguous()


def cmp(name, ours, ref, hw):
    """ours: [B*h*w, Cpad] ; ref: NCHW"""
    Bc, C, hh, ww = ref.shape
    assert (hh, ww) == tuple(hw), (name, hh, ww, hw)
    o = ours.float().view(Bc, hh, ww, -1)[..., :C].permute(0, 3, 1, 2)
    pad = ours.float().view(Bc, hh, ww, -1)[..., C:]
    ma, rel = O.parity(o, ref)
    print(f"{name:28s} shape {tuple(ref.shape)} rel {rel:.3e} max-abs {ma:.3e}  pad-abs-max {float(pad.abs().max()) if pad.numel() else 0:.2e}", flush=True)


toks = [nlc(f) for f in feats]
Wt = D.DPTWeights(m.dpt_feature_head1, m.dpt_regressor_head1)
tape = D.Tape()
p = "dpt_feature_head1."
maps, sizes, refs = [], [], []
for j in range(4):
    q = f"{p}input_process.{j}.0."
    r = F.conv2d(feats[j], sd[q + "0.weight"], sd[q + "0.bias"])
    x = D.conv1x1(tape, toks[j], Wt.pre[j])
    cmp(f"stage{j} conv1x1", x, r, (h, w))
    if j == 0 or j == 1:
        s = Wt.up[j].s
        r = F.conv_transpose2d(r, sd[q + "1.weight"], sd[q + "1.bias"], stride=s)
        x, hw = D.conv_transpose(tape, x, B, h, w, Wt.up[j]), (h * s, w * s)
        cmp(f"stage{j} convT s{s}", x, r, hw)
    elif j == 2:
        hw = (h, w)
    else:
        r = F.conv2d(r, sd[q + "1.weight"], sd[q + "1.bias"], stride=2, padding=1)
        x, hw = D.conv3x3(tape, x, B, h, w, Wt.up[j]), ops.conv_out_hw(h, w, 2)
        cmp(f"stage{j} conv3x3 s2", x, r, hw)
    r = F.conv2d(r, sd[f"{p}scratch.layer_rn.{j}.weight"], None, padding=1)
    x = D.conv3x3(tape, x, B, hw[0], hw[1], Wt.rn[j])
    cmp(f"stage{j} layer_rn", x, r, hw)
    maps.append(x); sizes.append(hw); refs.append(r)
l0, l1, l2, l3 = maps
r0, r1, r2, r3 = refs
rp4 = O._fusion(sd, p + "scratch.refinenet4.", r3, None)
p4 = D._fusion(tape, l3, None, B, sizes[3][0], sizes[3][1], Wt.fuse[3])
cmp("fusion4 (pre-crop)", p4, rp4, (2 * sizes[3][0], 2 * sizes[3][1]))
rp4 = rp4[:, :, : r2.shape[2], : r2.shape[3]]
p4 = D.crop(tape, p4, B, 2 * sizes[3][0], 2 * sizes[3][1], sizes[2][0], sizes[2][1])
cmp("fusion4 cropped", p4, rp4, sizes[2])
rp3 = O._fusion(sd, p + "scratch.refinenet3.", rp4, r2)
p3 = D._fusion(tape, p4, l2, B, sizes[2][0], sizes[2][1], Wt.fuse[2])
cmp("fusion3", p3, rp3, (2 * sizes[2][0], 2 * sizes[2][1]))
rp2 = O._fusion(sd, p + "scratch.refinenet2.", rp3, r1)
p2 = D._fusion(tape, p3, l1, B, sizes[1][0], sizes[1][1], Wt.fuse[1])
cmp("fusion2", p2, rp2, (2 * sizes[1][0], 2 * sizes[1][1]))
rp1 = O._fusion(sd, p + "scratch.refinenet1.", rp2, r0)
p1 = D._fusion(tape, p2, l0, B, sizes[0][0], sizes[0][1], Wt.fuse[0])
Hf, Wf = 2 * sizes[0][0], 2 * sizes[0][1]
cmp("fusion1", p1, rp1, (Hf, Wf))
q = "dpt_regressor_head1."
rc1 = F.conv2d(rp1, sd[q + "conv1.weight"], sd[q + "conv1.bias"], padding=1)
c1 = D.conv3x3(tape, p1, B, Hf, Wf, Wt.r1)
cmp("reg conv1", c1, rc1, (Hf, Wf))
ru = F.interpolate(rc1, size=(H, W), mode="bilinear", align_corners=True)
u = D.resize(tape, c1, B, Hf, Wf, H, W)
cmp("reg resize", u, ru, (H, W))
rc2 = F.relu(F.conv2d(ru, sd[q + "conv2.0.weight"], sd[q + "conv2.0.bias"], padding=1))
c2 = D.conv3x3(tape, u, B, H, W, Wt.r2, relu=True)
cmp("reg conv2+relu", c2, rc2, (H, W))
ro = F.conv2d(rc2, sd[q + "conv2.2.weight"], sd[q + "conv2.2.bias"])
o = D.conv1x1(tape, c2, Wt.r3, out_dtype=torch.float32)
cmp("reg out", o, ro, (H, W))
# engine's own decoder tokens vs oracle hooks
pk = m.pack(); pk.refresh_bf16()
tok, _ = m.encoder.forward_tokens(torch.cat((img1, img2), 0), pk, "encoder.")
half = tok.shape[0] // 2
t1, t2 = tok[:half], tok[half:]
take, _ = U.dust3r.feature_take_indices(m.info_sharing.depth, m.info_sharing.indices) if hasattr(U.dust3r, "feature_take_indices") else (list(cfg["ifr_indices"]), None)
(e1, e2), einter = m.info_sharing.forward_tokens([t1, t2], B, h, w, pk, "info_sharing.", take, False)
for name, ours, ref in (("hook0 enc", t1, feats[0]), ("hook1", einter[0][0], feats[1]), ("hook2", einter[1][0], feats[2]), ("hook3 final", e1, feats[3])):
    cmp(name, ours, ref, (h, w))
