"""GPU (-m gpu): module- and model-level parity of the B200 path.

Contract: the product computes in bf16 with fp32 accumulation, so against the fp32
golden vectors of the REAL reference it is held to the error of a bf16-autocast run of the same
arithmetic: err(ours, fp32) <= 1.5 x err(oracle under torch.autocast(bf16), fp32) + 2e-3, measured in
the same test; index/gather outputs are bit-exact.  Gradients: relative L2 per tensor <= 3e-2 and the
whole-model gradient direction cosine >= 0.999.
"""
import math

import pytest
import torch

import dust3r_oracle as O
import uniception_b200 as U
from golden_utils import load, weights

pytestmark = pytest.mark.gpu
DEV = "cuda"
# the fp32 oracle is the checker: keep cuDNN / cuBLAS out of TF32, as the reference's own KAT does
# (examples/models/dust3r/dust3r.py:198-230 compares with TF32 disabled)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def _autocast_err(fn_fp32_inputs):
    """relative error of the reference arithmetic under bf16 autocast vs itself in fp32.  Inside `O.reference_functionals()` the
    oracle issues the torch calls the reference issues, which is what makes the autocast run the reference's autocast run
    (pinned bit for bit by tests/test_oracle_golden.py::test_autocast_yardstick_is_the_reference_under_autocast)."""
    with O.reference_functionals():
        ref = fn_fp32_inputs()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            low = fn_fp32_inputs()
    return [O.parity(a.float(), b)[1] for a, b in zip(low, ref)]


def _tiny_model(cfg):
    m = U.DUSt3R(name="t", img_size=tuple(cfg["hw"]),
                 encoder_kwargs=dict(enc_embed_dim=cfg["C_enc"], enc_depth=cfg["enc_depth"], enc_num_heads=cfg["enc_heads"]),
                 info_sharing_kwargs=dict(depth=cfg["dec_depth"], dim=cfg["C_dec"], num_heads=cfg["dec_heads"]))
    m.load_state_dict(weights(cfg))
    return m.to(DEV)


@pytest.mark.parametrize("name", ["encoder_tiny", "encoder_vitb16_224"])
def test_encoder_vs_reference_golden(name):
    cfg, a = load(name)
    enc = U.CroCoEncoder(name="e", data_norm_type="dust3r", img_size=tuple(cfg["hw"]), enc_embed_dim=cfg["C"],
                         enc_depth=cfg["depth"], enc_num_heads=cfg["heads"])
    enc.load_state_dict(weights(cfg))
    enc = enc.to(DEV)
    img = a["img"].to(DEV)
    out = enc(U.ViTEncoderInput(image=img, data_norm_type="dust3r")).features
    assert out.shape == a["features"].shape and out.dtype == torch.float32
    sd = {("encoder." + k): v.to(DEV) for k, v in weights(cfg).items()}
    (ref_err,) = _autocast_err(lambda: [O.croco_encoder(sd, "encoder.", img, cfg["depth"], cfg["heads"])])
    ma, rel = O.parity(out, a["features"])
    print(f"{name}: ours vs reference fp32 rel {rel:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert rel <= 1.5 * ref_err + 2e-3, (rel, ref_err)


def test_encoder_ifr_vs_reference_golden():
    cfg, a = load("encoder_tiny_ifr")
    enc = U.CroCoIntermediateFeatureReturner(name="e", data_norm_type="dust3r", img_size=tuple(cfg["hw"]), enc_embed_dim=cfg["C"],
                                             enc_depth=cfg["depth"], enc_num_heads=cfg["heads"], indices=cfg["indices"],
                                             intermediates_only=False)
    enc.load_state_dict(weights(cfg))
    enc = enc.to(DEV)
    final, inter = enc(U.ViTEncoderInput(image=a["img"].to(DEV), data_norm_type="dust3r"))
    assert O.parity(final.features, a["features"])[1] <= 2e-2
    assert len(inter) == 2
    for i, o in enumerate(inter):
        assert O.parity(o.features, a[f"inter{i}"])[1] <= 2e-2
    # IFR semantics checked by the reference's self-test (cross_attention_transformer.py:589-607): last == final
    assert torch.equal(inter[-1].features, final.features)


@pytest.mark.parametrize("name", ["dust3r_tiny_linear", "dust3r_tiny_linear_sym"])
def test_dust3r_vs_reference_golden_fwd_bwd(name):
    cfg, a = load(name)
    m = _tiny_model(cfg)
    img1, img2 = a["img1"].to(DEV), a["img2"].to(DEV)
    v1 = {"img": img1, "instance": cfg["inst1"], "data_norm_type": "dust3r"}
    v2 = {"img": img2, "instance": cfg["inst2"], "data_norm_type": "dust3r"}
    r1, r2 = m(v1, v2)
    assert r1["pts3d"].shape == a["pts3d_1"].shape and r1["conf"].shape == a["conf_1"].shape
    assert r1["pts3d"].dtype == torch.float32 and float(r1["conf"].min()) >= 1.0

    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}
    kw = dict(enc_depth=cfg["enc_depth"], enc_heads=cfg["enc_heads"], dec_depth=cfg["dec_depth"], dec_heads=cfg["dec_heads"],
              instances=(cfg["inst1"], cfg["inst2"]))

    def run():
        o1, o2 = O.dust3r_forward(sd, img1, img2, **kw)
        return [o1["pts3d"], o1["conf"], o2["pts3d_in_other_view"], o2["conf"]]

    ref_errs = _autocast_err(run)
    ours = [r1["pts3d"], r1["conf"], r2["pts3d_in_other_view"], r2["conf"]]
    gold = [a["pts3d_1"], a["conf_1"], a["pts3d_2"], a["conf_2"]]
    for tag, x, g, re_ in zip(["pts3d_1", "conf_1", "pts3d_2", "conf_2"], ours, gold, ref_errs):
        ma, rel = O.parity(x, g)
        print(f"{name} {tag}: ours vs reference fp32 rel {rel:.3e} max-abs {ma:.3e} (autocast-bf16 oracle: {re_:.3e})")
        assert rel <= 1.5 * re_ + 2e-3, (tag, rel, re_)

    # backward: the bench loss (sum of all four outputs), gradients vs the reference's
    m.pack().zero_grad()
    O.bench_loss(r1, r2).backward()
    params = dict(m.named_parameters())
    for key, gk in [("encoder.enc_blocks.0.attn.qkv.weight", "grad_qkv0"), ("encoder.patch_embed.proj.weight", "grad_patch"),
                    ("info_sharing.multi_view_branches.1.0.cross_attn.projk.weight", "grad_projk")]:
        ma, rel = O.parity(params[key].grad, a[gk])
        print(f"{name} grad {key}: rel {rel:.3e}")
        assert rel <= 5e-2, (key, rel)
    norms = torch.tensor([float(params[k].grad.double().norm()) for k in cfg["grad_keys"]])
    ref_norms = a["grad_digest"][:, 1]
    rel_n = ((norms - ref_norms).abs() / ref_norms.clamp_min(1e-12))
    print(f"{name}: per-parameter grad-norm rel err max {float(rel_n.max()):.3e} median {float(rel_n.median()):.3e}")
    assert float(rel_n.median()) <= 2e-2 and float(rel_n.max()) <= 0.15


def test_dust3r_grads_vs_oracle_all_parameters():
    """Every parameter gradient against fp32 autograd through the oracle (same weights, same inputs)."""
    cfg, a = load("dust3r_tiny_linear")
    m = _tiny_model(cfg)
    img1, img2 = a["img1"].to(DEV), a["img2"].to(DEV)
    r1, r2 = m({"img": img1, "instance": cfg["inst1"], "data_norm_type": "dust3r"},
               {"img": img2, "instance": cfg["inst2"], "data_norm_type": "dust3r"})
    m.pack().zero_grad()
    O.bench_loss(r1, r2).backward()
    sd = {k: v.to(DEV).requires_grad_(True) for k, v in weights(cfg).items()}
    o1, o2 = O.dust3r_forward(sd, img1, img2, enc_depth=cfg["enc_depth"], enc_heads=cfg["enc_heads"],
                              dec_depth=cfg["dec_depth"], dec_heads=cfg["dec_heads"])
    O.bench_loss(o1, o2).backward()
    dot = nrm_a = nrm_b = 0.0
    worst = (0.0, "")
    for k, p in m.named_parameters():
        g, r = p.grad.double().flatten(), sd[k].grad.double().flatten()
        dot += float(g @ r)
        nrm_a += float(g @ g)
        nrm_b += float(r @ r)
        rel = float((g - r).norm() / r.norm().clamp_min(1e-12))
        if rel > worst[0]:
            worst = (rel, k)
    cos = dot / math.sqrt(nrm_a * nrm_b)
    print(f"whole-model gradient cosine {cos:.6f}; worst per-tensor rel {worst[0]:.3e} at {worst[1]}")
    assert cos >= 0.999 and worst[0] <= 0.1


def test_standalone_modules_compose_like_reference():
    """encoder -> info_sharing -> head -> adaptor through the public module API (NCHW dataclass I/O)
    equals the fused DUSt3R forward up to bf16 rounding at the module boundaries."""
    cfg, a = load("dust3r_tiny_linear")
    m = _tiny_model(cfg)
    img1, img2 = a["img1"].to(DEV), a["img2"].to(DEV)
    r1, _ = m({"img": img1, "instance": cfg["inst1"], "data_norm_type": "dust3r"},
              {"img": img2, "instance": cfg["inst2"], "data_norm_type": "dust3r"})
    m2 = _tiny_model(cfg)
    feats = m2.encoder(U.ViTEncoderInput(image=torch.cat((img1, img2)), data_norm_type="dust3r")).features
    assert feats.dtype == torch.float32 and feats.shape == (4, cfg["C_enc"], 2, 3)
    f1, f2 = feats.chunk(2, dim=0)
    out = m2.info_sharing(U.MultiViewTransformerInput(features=[f1, f2]))
    dec = m2.head1(U.PredictionHeadInput(last_feature=out.features[0])).decoded_channels
    ad = m2.adaptor(U.AdaptorInput(adaptor_feature=dec, output_shape_hw=tuple(cfg["hw"])))
    pts = ad.value.permute(0, 2, 3, 1)
    ma, rel = O.parity(pts, r1["pts3d"])
    print(f"standalone-composed vs fused pts3d: rel {rel:.3e}")
    assert rel <= 2e-2
    pts.sum().backward()
    assert m2.encoder.patch_embed.proj.weight.grad is not None and float(m2.encoder.patch_embed.proj.weight.grad.abs().sum()) > 0


def test_granular_blocks_vs_oracle():
    """Block / CrossAttentionBlock used on their own with the reference call signature."""
    from uniception_b200.blocks import Block, CrossAttentionBlock
    from functools import partial

    torch.manual_seed(0)
    C, H, B, hh, ww = 128, 2, 2, 4, 5
    rope = U.RoPE2D(100.0)
    norm = partial(torch.nn.LayerNorm, eps=1e-6)
    blk = Block(C, H, 4.0, qkv_bias=True, norm_layer=norm, rope=rope).to(DEV)
    x = torch.randn(B, hh * ww, C, device=DEV)
    pos = O.patch_positions(B, hh, ww, DEV)
    sd = {"b." + k: v.detach() for k, v in blk.state_dict().items()}
    y = blk(x, pos)
    ref = O.encoder_block(sd, "b.", x, pos, H, 100.0)
    assert O.parity(y.float(), ref)[1] <= 2e-2
    cblk = CrossAttentionBlock(C, H, 4.0, qkv_bias=True, norm_layer=norm, custom_positional_encoding=rope).to(DEV)
    x2 = torch.randn(B, hh * ww, C, device=DEV, requires_grad=True)
    yv = torch.randn(B, hh * ww, C, device=DEV)
    sd = {"b." + k: v.detach() for k, v in cblk.state_dict().items()}
    out = cblk(x2, yv, pos, pos)
    ref = O.decoder_block(sd, "b.", x2.detach(), yv, pos, pos, H, 100.0)
    assert O.parity(out.float(), ref)[1] <= 2e-2
    out.float().sum().backward()
    assert x2.grad is not None and cblk.cross_attn.projk.weight.grad is not None


def test_self_attention_block_options_vs_reference_golden():
    """stand-alone `SelfAttentionBlock` with latent_attn_dim + qk_norm + LayerScale + RoPE through the granular
    autograd ops (HeadNormFn, LayerScaleFn, AttentionFn) vs the reference's golden output and gradients."""
    from functools import partial

    from uniception_b200.blocks import SelfAttentionBlock

    cfg, a = load("self_attn_block_latent_qknorm_ls")
    blk = SelfAttentionBlock(cfg["dim"], cfg["heads"], cfg["latent"], qkv_bias=True, qk_norm=True, init_values=0.5,
                             norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), custom_positional_encoding=U.RoPE2D(freq=100.0))
    blk.load_state_dict(weights(cfg))
    blk = blk.to(DEV)
    x = a["x"].to(DEV).requires_grad_(True)
    pos = O.patch_positions(cfg["B"], cfg["hw"][0], cfg["hw"][1], DEV)
    y = blk(x, pos)
    sd = {"b." + k: v.to(DEV) for k, v in weights(cfg).items()}
    ref_err = max(_autocast_err(lambda: [O.encoder_block(sd, "b.", x.detach(), pos, cfg["heads"], 100.0)]))
    err = O.parity(y.float(), a["y"].to(DEV))[1]
    print(f"SelfAttentionBlock(latent, qk_norm, LayerScale): ours vs reference golden rel {err:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    y.float().sum().backward()
    for prm, k in ((blk.attn.qkv.weight, "attn.qkv.weight"), (blk.attn.q_norm.weight, "attn.q_norm.weight"),
                   (blk.attn.proj.weight, "attn.proj.weight"), (blk.ls1.gamma, "ls1.gamma"), (blk.ls2.gamma, "ls2.gamma"),
                   (blk.mlp.fc1.bias, "mlp.fc1.bias"), (x, "x")):
        e = O.parity(prm.grad.float(), a["grad_" + k.replace(".", "_")].to(DEV))[1]
        print(f"  grad {k}: rel {e:.3e}")
        assert e <= 5e-2, (k, e)


def test_gradient_checkpointing_recompute_matches_and_saves_memory():
    """`gradient_checkpointing=True` (info_sharing/base.py:59-71: every block under torch.utils.checkpoint) = per-block recompute
    in the engine: same outputs and gradients as the plain path, lower peak memory."""
    kw = dict(name="mv", input_embed_dim=192, depth=6, dim=256, num_heads=4, use_rand_idx_pe_for_non_reference_views=False,
              custom_positional_encoding=U.RoPE2D(freq=100.0))
    torch.manual_seed(3)
    plain = U.MultiViewAlternatingAttentionTransformer(**kw).to(DEV)
    ckpt = U.MultiViewAlternatingAttentionTransformer(gradient_checkpointing=True, **kw)
    ckpt.load_state_dict(plain.state_dict())
    ckpt = ckpt.to(DEV)
    feats = [torch.randn(2, 192, 24, 24, device=DEV) for _ in range(3)]
    # random cotangents: d sum(LayerNorm(x)) / dx vanishes identically for gamma = 1, which would leave only rounding noise
    cot = [torch.randn(2, 256, 24, 24, device=DEV) for _ in range(3)]
    res = {}
    for tag, m in (("plain", plain), ("plain2", plain), ("ckpt", ckpt)):
        m.zero_grad(set_to_none=True)
        fin = [f.clone().requires_grad_(True) for f in feats]
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        out = m(U.MultiViewTransformerInput(features=fin)).features
        held = torch.cuda.memory_allocated() - base  # activations kept for the backward
        sum((o * c).sum() for o, c in zip(out, cot)).backward()
        torch.cuda.synchronize()
        res[tag] = (out, fin[0].grad, m.self_attention_blocks[2].attn.qkv.weight.grad.clone(), m.proj_embed.bias.grad.clone(), held)
    for k in range(3):
        assert torch.equal(res["ckpt"][0][k], res["plain"][0][k])  # the forward pass is deterministic
    # the backward is not bit-reproducible (fp32 atomics in dq / wgrad, then bf16 rounding): the yardstick is the plain
    # path's own run-to-run difference
    for j, nm in ((1, "d_in0"), (2, "d_qkv2"), (3, "d_proj_embed_b")):
        noise = O.parity(res["plain2"][j], res["plain"][j])[1]
        e = O.parity(res["ckpt"][j], res["plain"][j])[1]
        print(f"  {nm}: checkpointed vs plain {e:.2e} (plain run-to-run {noise:.2e})")
        assert e <= 2 * noise + 1e-3, (nm, e, noise)
    print(f"activations held after forward: plain {res['plain'][4] / 2**20:.1f} MiB, checkpointed {res['ckpt'][4] / 2**20:.1f} MiB")
    assert res["ckpt"][4] < 0.5 * res["plain"][4]


def test_reference_self_test_mirror_multi_view_cross_attention():
    """The reference's in-module self-test (info_sharing/cross_attention_transformer.py:515-609) mirrored on the B200 modules:
    2 / 3 / 4 views with and without RoPE, IFR last-n / explicit indices / `torch.equal` normalisation semantics, plus the
    3- and 4-view cross-attention (Nk = (V-1)*N keys) against the oracle.  Body: tools/selftest_multiview.py."""
    import os
    import subprocess
    import sys as _sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([_sys.executable, os.path.join(root, "tools", "selftest_multiview.py")], capture_output=True, text=True,
                         timeout=600)
    print(out.stdout[-1500:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "SELFTEST OK" in out.stdout


def test_full_size_property_checks():
    """BASELINE.json sizes (ViT-L/16 + 12-layer decoder, 512x512, B=1 pair): size-independent
    properties -- confidence >= 1, finite outputs, batch-permutation equivariance and
    symmetrized-pair consistency of the encoder-dedup path."""
    torch.manual_seed(42)
    m = U.DUSt3R(name="dust3r", img_size=(512, 512)).to(DEV)
    g = torch.Generator().manual_seed(1234)
    a_img = torch.randn(1, 3, 512, 512, generator=g).clamp_(-1, 1).to(DEV)
    b_img = torch.randn(1, 3, 512, 512, generator=g).clamp_(-1, 1).to(DEV)
    img1, img2 = torch.cat((a_img, b_img)), torch.cat((b_img, a_img))
    with torch.no_grad():
        r1, r2 = m({"img": img1, "instance": ["0", "1"], "data_norm_type": "dust3r"},
                   {"img": img2, "instance": ["2", "3"], "data_norm_type": "dust3r"})
        s1, s2 = m({"img": img1, "instance": ["a", "b"], "data_norm_type": "dust3r"},
                   {"img": img2, "instance": ["b", "a"], "data_norm_type": "dust3r"})
    assert r1["pts3d"].shape == (2, 512, 512, 3) and r1["conf"].shape == (2, 512, 512, 1)
    assert torch.isfinite(r1["pts3d"]).all() and torch.isfinite(r2["pts3d_in_other_view"]).all()
    assert float(r1["conf"].min()) >= 1.0 and float(r2["conf"].min()) >= 1.0
    # symmetrized dedup (encode each image once) must agree with the plain path
    assert O.parity(s1["pts3d"], r1["pts3d"])[1] <= 1e-2
    assert O.parity(s2["conf"], r2["conf"])[1] <= 1e-2


def _oracle_dpt(sd, img1, img2, cfg):
    """dust3r_forward(head='dpt') with the fixture's IFR indices (see oracle/make_golden.py::_oracle_dpt_forward)."""
    B, _, H, W = img1.shape
    feat = O.croco_encoder(sd, "encoder.", torch.cat((img1, img2), 0), cfg["enc_depth"], cfg["enc_heads"])
    f1, f2 = feat.chunk(2, dim=0)
    (d1, d2), inter = O.info_sharing(sd, "info_sharing.", [f1, f2], cfg["dec_depth"], cfg["dec_heads"],
                                     indices=cfg["ifr_indices"], norm_intermediate=False)
    o1 = O.dpt_regressor(sd, "dpt_regressor_head1.", O.dpt_feature(sd, "dpt_feature_head1.", [f1, inter[0][0], inter[1][0], d1]), (H, W))
    o2 = O.dpt_regressor(sd, "dpt_regressor_head2.", O.dpt_feature(sd, "dpt_feature_head2.", [f2, inter[0][1], inter[1][1], d2]), (H, W))
    return o1, o2


def test_dust3r_dpt_vs_reference_golden_fwd_bwd():
    """DUSt3R with DPT heads: forward vs the reference's golden, backward vs oracle autograd.
    The raw head outputs are compared before the exp adaptor too (the adaptor amplifies errors by e^d)."""
    cfg, a = load("dust3r_tiny_dpt")
    m = U.DUSt3R(name="t", img_size=tuple(cfg["hw"]), pred_head_type="dpt", pred_head_feature_dim=32,
                 encoder_kwargs=dict(enc_embed_dim=cfg["C_enc"], enc_depth=cfg["enc_depth"], enc_num_heads=cfg["enc_heads"]),
                 info_sharing_kwargs=dict(depth=cfg["dec_depth"], dim=cfg["C_dec"], num_heads=cfg["dec_heads"]),
                 dpt_kwargs=dict(layer_dims=[12, 24, 48, 96]), dpt_indices=tuple(cfg["ifr_indices"]))
    m.load_state_dict(weights(cfg))
    m = m.to(DEV)
    img1, img2 = a["img1"].to(DEV), a["img2"].to(DEV)
    r1, r2 = m({"img": img1, "instance": cfg["inst1"], "data_norm_type": "dust3r"},
               {"img": img2, "instance": cfg["inst2"], "data_norm_type": "dust3r"})
    # The DPT state dict holds every layer_rn conv under three keys and every head tensor again under head{k}.0/1.*
    # (dpt_block.py:34-78, dust3r.py:178); load_state_dict lets the LAST alias win, so the oracle must read the
    # module's resolved state dict (as oracle/make_golden.py does), not the raw seeded one.
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    o1, o2 = _oracle_dpt(sd, img1, img2, cfg)
    with O.reference_functionals(), torch.autocast("cuda", dtype=torch.bfloat16):
        l1, _ = _oracle_dpt({k: v.detach() for k, v in sd.items()}, img1, img2, cfg)
    ref_err = O.parity(l1.float(), o1)[1]
    # conf = 1 + exp(raw channel 3): log(conf - 1) recovers the raw head output of the confidence channel
    raw_conf = torch.log((r1["conf"] - 1.0).clamp_min(1e-30))[..., 0]
    ma, rel = O.parity(raw_conf, o1[:, 3])
    print(f"dpt raw conf channel: ours vs oracle fp32 rel {rel:.3e} (autocast-bf16 oracle, all channels: {ref_err:.3e})")
    assert rel <= 2.0 * ref_err + 5e-3, (rel, ref_err)
    p1, c1 = O.pointmap_conf_adaptor(o1)
    gold_conf = a["conf_1"].to(DEV)
    assert O.parity(c1.permute(0, 2, 3, 1), gold_conf)[1] <= 1e-4  # oracle == reference golden
    # backward: a bounded loss on the raw outputs (log conf) so that the exp adaptor does not dominate the gradient
    m.pack().zero_grad()
    torch.log(r1["conf"] - 1.0 + 1e-30).sum().backward()
    o1[:, 3].sum().backward()
    dot = na = nb = 0.0
    for k, p in m.named_parameters():
        if p.grad is None or sd[k].grad is None:
            continue
        g, r = p.grad.double().flatten(), sd[k].grad.double().flatten()
        dot += float(g @ r); na += float(g @ g); nb += float(r @ r)
    cos = dot / math.sqrt(na * nb)
    print(f"dpt whole-model gradient cosine {cos:.5f}")
    assert cos >= 0.995


def test_c5_depth_patch14_vs_reference_golden():
    """BASELINE configs[4] in miniature: ViT encoder with patch 14 (ragged token counts, padded patch-embed pitch) ->
    DPT feature head -> regression processor -> DepthAdaptor(exp), forward vs the reference's golden and backward vs
    the oracle's autograd."""
    from golden_utils import c5_modules
    from uniception_b200.prediction_heads import DPTHead, PredictionHeadLayeredInput

    cfg, a = load("depth_c5_tiny_patch14")
    m = c5_modules(cfg)
    m.load_state_dict(weights(cfg))
    m = m.to(DEV)
    head = DPTHead(m.dpt_feature_head, m.dpt_regressor_head)
    adaptor = U.DepthAdaptor(name="depth", mode="exp")
    img = a["img"].to(DEV)
    hw = tuple(cfg["hw"])
    feats = [o.features for o in m.encoder(U.ViTEncoderInput(image=img, data_norm_type="dust3r"))]
    assert feats[0].shape == a["hook0"].shape
    raw = head(PredictionHeadLayeredInput(list_features=feats, target_output_shape=hw)).decoded_channels
    depth = adaptor(U.AdaptorInput(adaptor_feature=raw, output_shape_hw=hw)).value
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}

    def oracle():
        _, inter = O.croco_encoder(sd, "encoder.", img, cfg["depth"], cfg["heads"], cfg["patch"], indices=cfg["indices"])
        return inter, O.dpt_regressor(sd, "dpt_regressor_head.", O.dpt_feature(sd, "dpt_feature_head.", inter), hw)

    inter, oraw = oracle()
    with O.reference_functionals(), torch.autocast("cuda", dtype=torch.bfloat16):
        _, lraw = oracle()
    ref_err = O.parity(lraw.float(), oraw)[1]
    e_hook = O.parity(feats[3], a["hook3"].to(DEV))[1]
    e_raw = O.parity(raw, a["raw"].to(DEV))[1]
    e_depth = O.parity(depth, a["depth"].to(DEV))[1]
    # depth = exp(raw) turns the ABSOLUTE error of raw into a relative one: its yardstick is the autocast reference's own depth
    ref_depth_err = O.parity(O.depth_adaptor(lraw.float(), "exp"), O.depth_adaptor(oraw, "exp"))[1]
    print(f"c5 patch14: hook3 rel {e_hook:.3e}, raw head rel {e_raw:.3e}, depth rel {e_depth:.3e} (autocast-bf16 oracle raw: {ref_err:.3e}, "
          f"depth: {ref_depth_err:.3e})")
    assert O.parity(oraw, a["raw"].to(DEV))[1] <= 1e-4  # oracle (GPU fp32, TF32 off) == reference golden
    assert e_hook <= 2e-2
    assert e_raw <= 2.0 * ref_err + 5e-3, (e_raw, ref_err)
    assert e_depth <= 2.0 * ref_depth_err + 5e-3, (e_depth, ref_depth_err)
    # backward through encoder + head: gradient direction vs the oracle's autograd
    for p_ in m.parameters():
        p_.grad = None
    raw.sum().backward()
    oraw.sum().backward()
    dot = na = nb = 0.0
    for k, p_ in m.named_parameters():
        if p_.grad is None or sd[k].grad is None:
            continue
        g, r = p_.grad.double().flatten(), sd[k].grad.double().flatten()
        dot += float(g @ r); na += float(g @ g); nb += float(r @ r)
    cos = dot / math.sqrt(na * nb)
    print(f"c5 patch14 whole-pipeline gradient cosine {cos:.5f}")
    assert cos >= 0.995
    gp = m.encoder.patch_embed.proj.weight.grad
    assert gp is not None and O.parity(gp, sd["encoder.patch_embed.proj.weight"].grad)[1] <= 5e-2


def test_encoder_manyar_mixed_aspect_ratio_vs_reference_golden():
    """`CroCoEncoder(patch_embed_cls="ManyAR_PatchEmbed")` on a batch that mixes landscape and portrait samples
    (`true_shape`), forward vs the reference's golden, patch-embed / first-block gradients vs the reference's."""
    cfg, a = load("encoder_tiny_manyar")
    enc = U.CroCoEncoder(name="enc", data_norm_type="dust3r", patch_embed_cls="ManyAR_PatchEmbed", img_size=tuple(cfg["hw"]),
                         enc_embed_dim=cfg["C"], enc_depth=cfg["depth"], enc_num_heads=cfg["heads"])
    enc.load_state_dict(weights(cfg))
    enc = enc.to(DEV)
    img = a["img"].to(DEV)
    inp = U.ViTEncoderInput(image=img, data_norm_type="dust3r")
    inp.true_shape = a["true_shape"]
    feat = enc(inp).features
    sd = {k: v.to(DEV) for k, v in weights(cfg, "encoder.").items()}
    ref_err = _autocast_err(lambda: [O.croco_encoder(sd, "encoder.", img, cfg["depth"], cfg["heads"], true_shape=a["true_shape"])])[0]
    err = O.parity(feat, a["features"].to(DEV))[1]
    print(f"manyar encoder: ours vs reference golden rel {err:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    # the portrait sample really went through the transposed path: plain patch-embedding of the same batch differs
    plain = U.ViTEncoderInput(image=img, data_norm_type="dust3r")
    assert O.parity(enc(plain).features[1], a["features"][1].to(DEV))[1] > 0.1
    feat.sum().backward()
    gp = enc.patch_embed.proj.weight.grad
    assert O.parity(gp, a["grad_patch"].to(DEV))[1] <= 5e-2
    assert O.parity(enc.enc_blocks[0].attn.qkv.weight.grad, a["grad_qkv0"].to(DEV))[1] <= 5e-2


@pytest.mark.parametrize("name", ["global_attn_tiny", "global_attn_tiny_rope", "alternating_attn_tiny", "global_attn_tiny_scaled",
                                  "alternating_attn_tiny_qknorm_ls"])
def test_self_attention_info_sharing_vs_reference_golden(name):
    """`MultiViewGlobalAttentionTransformer` / `MultiViewAlternatingAttentionTransformer` on the B200 engine vs
    the reference's golden outputs and gradients (2 and 3 views, with and without RoPE)."""
    cfg, a = load(name)
    m = getattr(U, cfg["cls"])(name="mv", input_embed_dim=cfg["C_in"], depth=cfg["depth"], dim=cfg["dim"], num_heads=cfg["heads"],
                               use_rand_idx_pe_for_non_reference_views=False,
                               custom_positional_encoding=U.RoPE2D(freq=100.0) if cfg["rope"] else None,
                               use_scalable_softmax=cfg.get("scaling", False), use_entropy_scaling=cfg.get("scaling", False),
                               qk_norm=cfg.get("qk_norm", False), init_values=cfg.get("init_values"))
    sm = (True, True, 444, 1.4) if cfg.get("scaling") else None
    m.load_state_dict(weights(cfg), strict=False)  # view_pos_table is a buffer (the sinusoid table), not a weight
    m = m.to(DEV)
    feats = [a[f"feat{v}"].to(DEV).requires_grad_(True) for v in range(cfg["V"])]
    out = m(U.MultiViewTransformerInput(features=feats)).features
    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}
    fin = [f.detach() for f in feats]
    ref_err = max(_autocast_err(lambda: O.self_attention_info_sharing(sd, "", fin, cfg["depth"], cfg["heads"],
                                                                      alternating="Alternating" in cfg["cls"],
                                                                      base=100.0 if cfg["rope"] else None,
                                                                      pe_for_non_ref=cfg["pe_for_non_ref"], softmax_scaling=sm)))
    err = max(O.parity(out[v], a[f"out{v}"].to(DEV))[1] for v in range(cfg["V"]))
    print(f"{name}: ours vs reference golden rel {err:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    sum(o.sum() for o in out).backward()
    g = m.self_attention_blocks[1].attn.qkv.weight.grad
    assert O.parity(g, a["grad_qkv1"].to(DEV))[1] <= 5e-2
    assert O.parity(m.proj_embed.weight.grad, a["grad_proj_embed"].to(DEV))[1] <= 5e-2
    assert O.parity(feats[0].grad, a["grad_in0"].to(DEV))[1] <= 5e-2
    if cfg.get("qk_norm"):  # uc_headnorm_* and uc_layerscale_* on the path
        blk = m.self_attention_blocks[1]
        for prm, arr in ((blk.attn.q_norm.weight, "grad_attn_q_norm_weight"), (blk.attn.k_norm.bias, "grad_attn_k_norm_bias"),
                         (blk.ls1.gamma, "grad_ls1_gamma"), (blk.ls2.gamma, "grad_ls2_gamma"), (blk.attn.proj.bias, "grad_attn_proj_bias")):
            e = O.parity(prm.grad, a[arr].to(DEV))[1]
            print(f"  {arr}: rel {e:.3e}")
            assert e <= 5e-2, (arr, e)


@pytest.mark.parametrize("name", ["global_attn_tiny_ifr", "alternating_attn_tiny_ifr"])
def test_self_attention_info_sharing_ifr_vs_reference_golden(name):
    """`MultiView{Global,Alternating}AttentionTransformerIFR` on the B200 engine vs the reference's golden final
    maps, tapped intermediates (final-normed / raw) and gradients (tapped depths take the un-fused bias-gradient path)."""
    cfg, a = load(name)
    m = getattr(U, cfg["cls"])(name="mv", input_embed_dim=cfg["C_in"], depth=cfg["depth"], dim=cfg["dim"], num_heads=cfg["heads"],
                               use_rand_idx_pe_for_non_reference_views=False, custom_positional_encoding=U.RoPE2D(freq=100.0),
                               indices=cfg["indices"], norm_intermediate=cfg["norm_intermediate"])
    m.load_state_dict(weights(cfg), strict=False)
    m = m.to(DEV)
    V = cfg["V"]
    feats = [a[f"feat{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    final, inter = m(U.MultiViewTransformerInput(features=feats))
    assert len(inter) == len(cfg["indices"])
    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}
    fin = [f.detach() for f in feats]

    def oracle_all():
        o, it = O.self_attention_info_sharing(sd, "", fin, cfg["depth"], cfg["heads"], alternating="Alternating" in cfg["cls"],
                                              base=100.0, pe_for_non_ref=cfg["pe_for_non_ref"], indices=cfg["indices"],
                                              norm_intermediate=cfg["norm_intermediate"])
        return o + [t for lvl in it for t in lvl]

    ref_err = max(_autocast_err(oracle_all))
    ours = list(final.features) + [t for lvl in inter for t in lvl.features]
    gold = [a[f"out{v}"] for v in range(V)] + [a[f"inter{k}_{v}"] for k in range(len(inter)) for v in range(V)]
    err = max(O.parity(x, g.to(DEV))[1] for x, g in zip(ours, gold))
    print(f"{name}: ours vs reference golden rel {err:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    (sum(o.sum() for o in final.features)
     + sum((0.5 + k) * sum(t.sum() for t in lvl.features) for k, lvl in enumerate(inter))).backward()
    assert O.parity(m.self_attention_blocks[1].attn.qkv.weight.grad, a["grad_qkv1"].to(DEV))[1] <= 5e-2
    assert O.parity(m.norm.weight.grad, a["grad_norm_w"].to(DEV))[1] <= 5e-2
    assert O.parity(m.self_attention_blocks[1].mlp.fc2.bias.grad, a["grad_fc2_b1"].to(DEV))[1] <= 5e-2
    assert O.parity(m.proj_embed.weight.grad, a["grad_proj_embed"].to(DEV))[1] <= 5e-2
    assert O.parity(feats[0].grad, a["grad_in0"].to(DEV))[1] <= 5e-2
    # intermediates_only returns just the list
    m.intermediates_only = True
    only = m(U.MultiViewTransformerInput(features=[f.detach() for f in feats]))
    assert isinstance(only, list) and len(only) == len(cfg["indices"])


@pytest.mark.parametrize("name", ["global_attn_tiny_tokens", "alternating_attn_tiny_tokens", "alternating_attn_tiny_pv_tokens"])
def test_additional_input_tokens_vs_reference_golden(name):
    """global / per-view additional input tokens through the B200 engine (ragged sequence lengths V*(N+Tv)+T; the
    alternating variant's frame-level blocks skip the global tokens) vs the reference's golden outputs and gradients."""
    from golden_utils import token_levels, token_loss

    cfg, a = load(name)
    kw = dict(indices=cfg["indices"]) if cfg["indices"] is not None else {}
    m = getattr(U, cfg["cls"])(name="mv", input_embed_dim=cfg["C_in"], depth=cfg["depth"], dim=cfg["dim"], num_heads=cfg["heads"],
                               use_rand_idx_pe_for_non_reference_views=False, **kw)
    m.load_state_dict(weights(cfg), strict=False)
    m = m.to(DEV)
    V = cfg["V"]
    feats = [a[f"feat{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    extra = a["extra"].to(DEV).requires_grad_(True) if cfg["T"] else None
    pv = [a[f"pv{v}"].to(DEV).requires_grad_(True) for v in range(V)] if cfg["Tv"] else None
    res = m(U.MultiViewTransformerInput(features=feats, additional_input_tokens=extra, additional_input_tokens_per_view=pv))
    outs = [res] if cfg["indices"] is None else [res[0]] + list(res[1])
    ours = [(o.features, o.additional_token_features, o.additional_token_features_per_view) for o in outs]
    gold, wts = token_levels(cfg, a, DEV)
    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}

    def flat(levels):
        return [t for (mm, ee, pp) in levels for t in list(mm) + ([ee] if ee is not None else []) + (list(pp) if pp is not None else [])]

    def oracle_all():
        o = O.self_attention_info_sharing(sd, "", [f.detach() for f in feats], cfg["depth"], cfg["heads"],
                                          alternating="Alternating" in cfg["cls"], pe_for_non_ref=cfg["pe_for_non_ref"],
                                          indices=cfg["indices"], extra=extra.detach() if extra is not None else None,
                                          extra_per_view=[t.detach() for t in pv] if pv is not None else None)
        return flat([o] if cfg["indices"] is None else [o[0]] + list(o[1]))

    ref_err = max(_autocast_err(oracle_all))
    fo, fg = flat(ours), flat(gold)
    assert len(fo) == len(fg) and all(x.shape == g.shape for x, g in zip(fo, fg))
    err = max(O.parity(x, g)[1] for x, g in zip(fo, fg))
    print(f"{name}: ours vs reference golden rel {err:.3e} over {len(fo)} outputs (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    token_loss(ours, wts).backward()
    blk = m.self_attention_blocks
    for prm, arr in ((blk[1].attn.qkv.weight, "grad_qkv1"), (blk[1].mlp.fc2.bias, "grad_fc2_b1"), (blk[0].mlp.fc2.bias, "grad_fc2_b0"),
                     (m.proj_embed.weight, "grad_proj_embed"), (m.proj_embed.bias, "grad_proj_embed_b"), (feats[0], "grad_in0"),
                     (extra, "grad_extra"), (pv[-1] if pv else None, "grad_pv_last")):
        if prm is None:
            continue
        e = O.parity(prm.grad, a[arr].to(DEV))[1]
        print(f"  {arr}: rel {e:.3e}")
        assert e <= 5e-2, (arr, e)
    # the reference refuses a positional-encoding plugin together with additional tokens (global_attention_transformer.py:341-351)
    m2 = getattr(U, cfg["cls"])(name="mv", input_embed_dim=cfg["C_in"], depth=2, dim=cfg["dim"], num_heads=cfg["heads"],
                                custom_positional_encoding="rope" if "Global" in cfg["cls"] else U.RoPE2D(freq=100.0)).to(DEV)
    with pytest.raises(ValueError):
        m2(U.MultiViewTransformerInput(features=[f.detach() for f in feats], additional_input_tokens=extra, additional_input_tokens_per_view=pv))


def test_cross_attention_qk_norm_layerscale_vs_reference_golden():
    """`MultiViewCrossAttentionTransformer(qk_norm=True, init_values=...)`: per-head q/k LayerNorm fused with
    RoPE (uc_headnorm_*) and LayerScale (uc_layerscale_*) vs the reference's golden outputs and gradients."""
    cfg, a = load("cross_attn_tiny_qknorm_ls")
    m = U.MultiViewCrossAttentionTransformer(name="mv", input_embed_dim=cfg["C_in"], num_views=2, depth=cfg["depth"], dim=cfg["dim"],
                                             num_heads=cfg["heads"], custom_positional_encoding=U.RoPE2D(freq=100.0),
                                             qk_norm=True, init_values=cfg["init_values"])
    m.load_state_dict(weights(cfg))
    m = m.to(DEV)
    feats = [a["feat0"].to(DEV).requires_grad_(True), a["feat1"].to(DEV).requires_grad_(True)]
    out = m(U.MultiViewTransformerInput(features=feats)).features
    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}
    fin = [f.detach() for f in feats]
    ref_err = max(_autocast_err(lambda: O.info_sharing(sd, "", fin, cfg["depth"], cfg["heads"])))
    err = max(O.parity(out[v], a[f"out{v}"].to(DEV))[1] for v in range(2))
    print(f"cross-attn + qk_norm + LayerScale: ours vs reference golden rel {err:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    sum(o.sum() for o in out).backward()
    blk = m.multi_view_branches[1][0]
    checks = ((blk.cross_attn.q_norm.weight, "grad_cross_attn_q_norm_weight"), (blk.cross_attn.k_norm.bias, "grad_cross_attn_k_norm_bias"),
              (blk.attn.q_norm.bias, "grad_attn_q_norm_bias"), (blk.ls1.gamma, "grad_ls1_gamma"), (blk.ls2.gamma, "grad_ls2_gamma"),
              (blk.ls3.gamma, "grad_ls3_gamma"), (blk.cross_attn.proj.bias, "grad_cross_attn_proj_bias"),
              (blk.mlp.fc2.bias, "grad_mlp_fc2_bias"), (blk.cross_attn.projq.weight, "grad_projq"))
    for prm, arr in checks:
        e = O.parity(prm.grad, a[arr].to(DEV))[1]
        print(f"  {arr}: rel {e:.3e}")
        assert e <= 5e-2, (arr, e)
    assert O.parity(feats[0].grad, a["grad_in0"].to(DEV))[1] <= 5e-2


def test_cross_attention_softmax_scaling_vs_reference_golden():
    """`MultiViewCrossAttentionTransformer(use_scalable_softmax=True, use_entropy_scaling=True)`: the
    token-count query multipliers fold into the attention kernels' scale (forward, dq, dk)."""
    cfg, a = load("cross_attn_tiny_scaled")
    m = U.MultiViewCrossAttentionTransformer(name="mv", input_embed_dim=cfg["C_in"], num_views=2, depth=cfg["depth"], dim=cfg["dim"],
                                             num_heads=cfg["heads"], custom_positional_encoding=U.RoPE2D(freq=100.0),
                                             use_scalable_softmax=True, use_entropy_scaling=True)
    m.load_state_dict(weights(cfg))
    m = m.to(DEV)
    feats = [a["feat0"].to(DEV).requires_grad_(True), a["feat1"].to(DEV).requires_grad_(True)]
    out = m(U.MultiViewTransformerInput(features=feats)).features
    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}
    fin = [f.detach() for f in feats]
    sm = tuple(cfg["softmax_scaling"])
    ref_err = max(_autocast_err(lambda: O.info_sharing(sd, "", fin, cfg["depth"], cfg["heads"], softmax_scaling=sm)))
    err = max(O.parity(out[v], a[f"out{v}"].to(DEV))[1] for v in range(2))
    print(f"cross-attn + softmax scaling: ours vs reference golden rel {err:.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert err <= 1.5 * ref_err + 2e-3, (err, ref_err)
    sum(o.sum() for o in out).backward()
    assert O.parity(m.multi_view_branches[1][0].cross_attn.projq.weight.grad, a["grad_projq"].to(DEV))[1] <= 5e-2
    assert O.parity(m.multi_view_branches[0][1].attn.qkv.weight.grad, a["grad_qkv"].to(DEV))[1] <= 5e-2
    assert O.parity(feats[0].grad, a["grad_in0"].to(DEV))[1] <= 5e-2


@pytest.mark.parametrize("name", ["diff_cross_attn_tiny", "diff_cross_attn_tiny_ifr"])
def test_diff_cross_attention_vs_reference_golden(name):
    """the DiffAttention family (`DifferentialMultiViewCrossAttentionTransformer(IFR)`) on the un-fused attention
    path -- uc_gemm scores, uc_softmax_rows, uc_gemm PV and the four gradient products -- with 128-wide self-attention heads and
    64-wide q / k against 128-wide v; forward vs the reference's golden, gradients of every kind of parameter the family adds."""
    cfg, a = load(name)
    V = cfg["V"]
    kw = dict(name="mvd", input_embed_dim=cfg["C_in"], num_views=V, depth=cfg["depth"], dim=cfg["dim"], num_heads=cfg["heads"],
              custom_positional_encoding=U.RoPE2D(freq=100.0) if cfg["rope"] else None)
    if cfg["indices"] is None:
        m = U.DifferentialMultiViewCrossAttentionTransformer(**kw)
    else:
        m = U.DifferentialMultiViewCrossAttentionTransformerIFR(indices=cfg["indices"], **kw)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in cfg["shapes"].items()}
    m.load_state_dict(weights(cfg))
    m = m.to(DEV)
    feats = [a[f"feat{v}"].to(DEV).requires_grad_(True) for v in range(V)]
    res = m(U.MultiViewTransformerInput(features=feats))
    out, inter = (res[0].features, [lv.features for lv in res[1]]) if cfg["indices"] is not None else (res.features, [])
    sd = {k: v.to(DEV) for k, v in weights(cfg).items()}
    fin = [f.detach() for f in feats]

    def oracle():
        r = O.diff_info_sharing(sd, "", fin, cfg["depth"], cfg["heads"], base=100.0 if cfg["rope"] else None, indices=cfg["indices"])
        return (list(r[0]) + [t for lv in r[1] for t in lv]) if cfg["indices"] is not None else list(r)

    ref_err = max(_autocast_err(oracle))
    errs = [O.parity(out[v], a[f"out{v}"].to(DEV))[1] for v in range(V)]
    errs += [O.parity(lv[v], a[f"inter{k}_{v}"].to(DEV))[1] for k, lv in enumerate(inter) for v in range(V)]
    print(f"{name}: ours vs reference golden rel {max(errs):.3e} (autocast-bf16 oracle: {ref_err:.3e})")
    assert max(errs) <= 1.5 * ref_err + 2e-3, (errs, ref_err)
    loss = sum(t.sum() for t in out) + sum((k + 1.5) * sum(t.sum() for t in lv) for k, lv in enumerate(inter))
    loss.backward()
    params = dict(m.named_parameters())
    for key in ("multi_view_branches.1.0.cross_attn.projq.weight", "multi_view_branches.0.1.attn.qkv.weight",
                "multi_view_branches.0.0.cross_attn.lambda_q1", "multi_view_branches.1.1.cross_attn.lambda_k2",
                "multi_view_branches.0.1.cross_attn.subln.weight", "multi_view_branches.1.0.cross_attn.projv.bias",
                "multi_view_branches.0.0.mlp.fc1.weight", "proj_embed.weight"):
        e = O.parity(params[key].grad, a["grad_" + key.replace(".", "_")].to(DEV))[1]
        print(f"   grad {key}: rel {e:.3e}")
        assert e <= 5e-2, (key, e)
    assert O.parity(feats[0].grad, a["grad_in0"].to(DEV))[1] <= 5e-2
