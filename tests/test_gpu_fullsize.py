"""GPU (-m gpu): parity AT SIZE -- the real BASELINE.json models, one image pair, forward + backward, against the fp32 oracle
run on the same GPU (TF32 off).

  * configs[1] / [2]: DUSt3R ViT-L/16 + 12-layer two-view decoder + linear heads at 224^2 and 512^2;
  * DUSt3R ViT-L/16 + DPT heads at 512^2 (a13-a15 inside the model);
  * configs[4]: ViT-L/14 intermediate-feature encoder -> DPTFeature -> DPTRegressionProcessor -> DepthAdaptor at 518^2
    (1369 ragged tokens), through the STAND-ALONE module API.

Yardstick, measured in the same test: the error of the reference arithmetic under
torch.autocast(bf16) -- the fused torch functionals the reference really calls (`O.reference_functionals()`), heads in fp32
with autocast disabled (factory/dust3r.py:285-309) -- against the same arithmetic in fp32.  Bar for every output tensor:
    err(ours) <= 1.0 x err(autocast reference) + 1e-3     (relative L2)
Gradients: every parameter gradient is compared; the per-stage relative L2 error (encoder / decoder / heads, all tensors
of the stage stacked) is held to the same 1.0 x + 1e-3 bar against the autocast reference's per-stage gradient error, and the
report names the worst tensor of each stage.
"""
import math

import pytest
import torch

import dust3r_oracle as O
import uniception_b200 as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

BAR_MULT, BAR_ABS = 1.0, 1e-3


def _images(S, B=1):
    g = torch.Generator().manual_seed(1234)
    a = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1)
    b = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1)
    return a.to(DEV), b.to(DEV)


def _stage_of(name: str) -> str:
    if name.startswith("encoder."):
        return "encoder"
    if name.startswith("info_sharing."):
        return "decoder"
    return "heads"


def _grad_report(tag, ours, ref, low):
    """ours / ref / low: dict name -> gradient (ours, fp32 oracle, autocast oracle).  Returns {stage: (err_ours, err_low)}."""
    acc = {}
    worst = {}
    for k, r in ref.items():
        if r is None or k not in ours or ours[k] is None:
            continue
        st = _stage_of(k)
        r64, o64, l64 = r.double().flatten(), ours[k].double().flatten(), low[k].double().flatten()
        a = acc.setdefault(st, [0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += float((o64 - r64).pow(2).sum())
        a[1] += float((l64 - r64).pow(2).sum())
        a[2] += float(r64.pow(2).sum())
        a[3] += float(o64 @ r64)
        a[4] += float(o64 @ o64)
        rel = float((o64 - r64).norm() / r64.norm().clamp_min(1e-30))
        rel_l = float((l64 - r64).norm() / r64.norm().clamp_min(1e-30))
        if st not in worst or rel - rel_l > worst[st][0] - worst[st][1]:
            worst[st] = (rel, rel_l, k)
    out = {}
    for st, (eo, el, rr, dot, oo) in acc.items():
        e_o, e_l = math.sqrt(eo / rr), math.sqrt(el / rr)
        cos = dot / math.sqrt(oo * rr)
        out[st] = (e_o, e_l)
        w = worst[st]
        print(f"{tag} grads[{st}]: ours {e_o:.3e} vs autocast-reference {e_l:.3e} (cosine {cos:.6f}); "
              f"largest excess: {w[2]} ours {w[0]:.3e} / autocast {w[1]:.3e}")
    return out


def _check(tag, name, e_ours, e_low, failures):
    ok = e_ours <= BAR_MULT * e_low + BAR_ABS
    print(f"{tag} {name}: ours {e_ours:.3e} vs autocast-reference {e_low:.3e} -> {'OK' if ok else 'EXCEEDS'} "
          f"(bar {BAR_MULT:.1f}x + {BAR_ABS:g})")
    if not ok:
        failures.append((name, e_ours, e_low))


def _oracle_runs(run, sd_src):
    """fp32 and autocast-bf16 runs of `run(sd) -> (outputs list, loss)`; returns (outs32, grads32, outs16, grads16)."""
    res = []
    for low in (False, True):
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd_src.items()}
        with O.reference_functionals():
            if low:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    outs, loss = run(sd)
            else:
                outs, loss = run(sd)
        loss.backward()
        res.append(([o.detach().float() for o in outs], {k: v.grad for k, v in sd.items()}))
        del loss, outs
    return res[0][0], res[0][1], res[1][0], res[1][1]


@pytest.mark.parametrize("S,head", [(224, "linear"), (512, "linear"), (512, "dpt")])
def test_dust3r_full_size_vs_oracle_fwd_bwd(S, head):
    tag = f"DUSt3R-L/16 {head} {S}^2"
    torch.manual_seed(42)
    m = U.DUSt3R(name="dust3r", img_size=(S, S), pred_head_type=head).to(DEV)
    img1, img2 = _images(S)
    v1 = {"img": img1, "instance": ["0"], "data_norm_type": "dust3r"}
    v2 = {"img": img2, "instance": ["1"], "data_norm_type": "dust3r"}
    m.zero_grad(set_to_none=True)  # the default of torch optimisers
    r1, r2 = m(v1, v2)
    O.bench_loss(r1, r2).backward()
    ours_out = [t.detach() for t in (r1["pts3d"], r1["conf"], r2["pts3d_in_other_view"], r2["conf"])]
    ours_g = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_src = {k: v.detach() for k, v in m.state_dict().items()}  # alias-resolved (DPT keys), see test_gpu_model.py

    def run(sd):
        o1, o2 = O.dust3r_forward(sd, img1, img2, head=head)
        return [o1["pts3d"], o1["conf"], o2["pts3d_in_other_view"], o2["conf"]], O.bench_loss(o1, o2)

    out32, g32, out16, g16 = _oracle_runs(run, sd_src)
    failures = []
    for name, x, r, l in zip(["pts3d_1", "conf_1", "pts3d_2", "conf_2"], ours_out, out32, out16):
        assert torch.isfinite(x).all()
        _check(tag, name, O.parity(x, r)[1], O.parity(l, r)[1], failures)
    # DPT state dicts alias tensors under several keys: compare through the parameter names only
    names = [k for k, _ in m.named_parameters()]
    st = _grad_report(tag, ours_g, {k: g32[k] for k in names}, {k: g16[k] for k in names})
    for stage, (e_o, e_l) in st.items():
        _check(tag, f"grads[{stage}]", e_o, e_l, failures)
    assert not failures, failures


def test_c5_vitl14_dpt_depth_518_standalone_modules_vs_oracle():
    """BASELINE configs[4] at size through the stand-alone module API: CroCoIntermediateFeatureReturner(patch 14) ->
    DPTFeature.forward -> DPTRegressionProcessor.forward -> DepthAdaptor, forward + backward."""
    from uniception_b200.prediction_heads import DPTFeature, DPTRegressionProcessor, PredictionHeadLayeredInput

    tag = "C5 ViT-L/14+DPT 518^2"
    S, patch, C, depth, heads, idx = 518, 14, 1024, 24, 16, [5, 11, 17, 23]
    torch.manual_seed(42)
    m = torch.nn.Module()
    m.encoder = U.CroCoIntermediateFeatureReturner(name="enc", data_norm_type="dust3r", img_size=(S, S), patch_size=patch,
                                                   enc_embed_dim=C, enc_depth=depth, enc_num_heads=heads, indices=idx,
                                                   intermediates_only=True)
    m.dpt_feature_head = DPTFeature(patch_size=patch, hooks=[0, 1, 2, 3], input_feature_dims=[C] * 4)
    m.dpt_regressor_head = DPTRegressionProcessor(input_feature_dim=256, output_dim=1)
    m = m.to(DEV)
    adaptor = U.DepthAdaptor(name="depth", mode="exp")
    img, _ = _images(S)
    feats = [o.features for o in m.encoder(U.ViTEncoderInput(image=img, data_norm_type="dust3r"))]
    assert feats[0].shape == (1, C, 37, 37)
    fin = m.dpt_feature_head(PredictionHeadLayeredInput(list_features=feats, target_output_shape=(S, S)))
    assert fin.features_upsampled_8x.shape == (1, 256, 296, 296) and fin.features_upsampled_8x.dtype == torch.float32
    raw = m.dpt_regressor_head(fin).decoded_channels
    assert raw.shape == (1, 1, S, S)
    depth_out = adaptor(U.AdaptorInput(adaptor_feature=raw, output_shape_hw=(S, S))).value
    raw.sum().backward()
    ours_g = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_src = {k: v.detach() for k, v in m.state_dict().items()}

    def run(sd):
        _, inter = O.croco_encoder(sd, "encoder.", img, depth, heads, patch, indices=idx)
        with torch.autocast("cuda", enabled=False):  # heads run in fp32 in the reference's models (factory/dust3r.py:309)
            f8 = O.dpt_feature(sd, "dpt_feature_head.", [t.float() for t in inter])
            r = O.dpt_regressor(sd, "dpt_regressor_head.", f8, (S, S))
        return [inter[3], f8, r, O.depth_adaptor(r.float(), "exp")], r.sum()

    out32, g32, out16, g16 = _oracle_runs(run, sd_src)
    failures = []
    ours = [feats[3].detach(), fin.features_upsampled_8x.detach(), raw.detach(), depth_out.detach()]
    for name, x, r, l in zip(["encoder hook 23", "features_upsampled_8x", "raw depth head", "depth"], ours, out32, out16):
        assert torch.isfinite(x).all()
        _check(tag, name, O.parity(x, r)[1], O.parity(l, r)[1], failures)
    names = [k for k, _ in m.named_parameters()]
    st = _grad_report(tag, ours_g, {k: g32[k] for k in names}, {k: g16[k] for k in names})
    for stage, (e_o, e_l) in st.items():
        _check(tag, f"grads[{stage}]", e_o, e_l, failures)
    assert not failures, failures


def test_dpt_heads_survive_zero_grad_set_to_none_over_two_steps():
    """with `zero_grad(set_to_none=True)` the DPT head gradients used to be discarded by the first
    ParamPack re-bind of the step.  Two steps, default zero_grad, a frozen encoder parameter in the mix: the DPT gradients
    must be non-zero, live in the flat buffer (what the optimiser / GradSync see) and be equal in both steps."""
    torch.manual_seed(0)
    m = U.DUSt3R(name="t", img_size=(32, 48), pred_head_type="dpt", pred_head_feature_dim=32,
                 encoder_kwargs=dict(enc_embed_dim=128, enc_depth=2, enc_num_heads=2),
                 info_sharing_kwargs=dict(depth=9, dim=128, num_heads=2), dpt_kwargs=dict(layer_dims=[12, 24, 48, 96])).to(DEV)
    m.encoder.patch_embed.proj.weight.requires_grad_(False)  # the old sentinel parameter, frozen
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.0)
    g = torch.Generator().manual_seed(1)
    img1 = torch.randn(2, 3, 32, 48, generator=g).to(DEV)
    img2 = torch.randn(2, 3, 32, 48, generator=g).to(DEV)
    v1 = {"img": img1, "instance": ["0", "1"], "data_norm_type": "dust3r"}
    v2 = {"img": img2, "instance": ["2", "3"], "data_norm_type": "dust3r"}
    snaps = []
    for step in range(2):
        opt.zero_grad()  # set_to_none=True
        r1, r2 = m(v1, v2)
        torch.log(r1["conf"]).sum().add(torch.log(r2["conf"]).sum()).backward()
        pk = m.pack()
        w = m.dpt_regressor_head1.conv1.weight
        assert w.grad is not None and float(w.grad.abs().sum()) > 0
        assert w.grad.data_ptr() == pk.grad_view("dpt_regressor_head1.conv1.weight").data_ptr()  # the flat-buffer view
        q = m.info_sharing.multi_view_branches[0][0].attn.qkv.weight
        assert q.grad is not None and float(q.grad.abs().sum()) > 0
        snaps.append((w.grad.clone(), q.grad.clone()))
        opt.step()
    # no accumulation across steps (the dropped gradients restarted from zero), up to the atomics' summation order
    assert O.parity(snaps[1][0], snaps[0][0])[1] <= 2e-2
    assert O.parity(snaps[1][1], snaps[0][1])[1] <= 2e-2


def test_standalone_dpt_modules_match_fused_head():
    """DPTFeature.forward -> DPTRegressionProcessor.forward (prediction_heads/dpt.py:180-232, :285-311) == DPTHead (the
    reference's nn.Sequential of both) on the same weights, forward and input gradients; hooks are honoured."""
    from uniception_b200.prediction_heads import DPTFeature, DPTHead, DPTRegressionProcessor, PredictionHeadLayeredInput

    torch.manual_seed(5)
    f = DPTFeature(patch_size=16, hooks=[3, 0, 2, 1], input_feature_dims=[128, 192, 192, 192], layer_dims=[24, 48, 96, 192],
                   feature_dim=64).to(DEV)
    r = DPTRegressionProcessor(input_feature_dim=64, output_dim=4).to(DEV)
    feats = [torch.randn(2, c, 6, 5, device=DEV, requires_grad=True) for c in (192, 192, 192, 128)]  # list index 3 -> hook 0
    inp = PredictionHeadLayeredInput(list_features=feats, target_output_shape=(96, 80))
    mid = f(inp)
    assert mid.features_upsampled_8x.shape == (2, 64, 48, 40) and mid.target_output_shape == (96, 80)
    out = r(mid).decoded_channels
    assert out.shape == (2, 4, 96, 80) and out.dtype == torch.float32
    cot = torch.randn_like(out)
    (out * cot).sum().backward()
    g_split = [t.grad.clone() for t in feats]
    gw_split = r.conv1.weight.grad.clone()
    for t in feats:
        t.grad = None
    r.zero_grad(set_to_none=True)
    f.zero_grad(set_to_none=True)
    ordered = [feats[h] for h in (3, 0, 2, 1)]
    out2 = DPTHead(f, r)(PredictionHeadLayeredInput(list_features=ordered, target_output_shape=(96, 80))).decoded_channels
    (out2 * cot).sum().backward()
    # the split path rounds the 8x feature map to fp32 NCHW and back to bf16 (exact) -- same kernels otherwise
    assert O.parity(out, out2)[1] <= 1e-5
    for a_, b_ in zip(g_split, [t.grad for t in feats]):
        assert O.parity(a_, b_)[1] <= 2e-2
    assert O.parity(gw_split, r.conv1.weight.grad)[1] <= 2e-2
    with pytest.raises(AssertionError):  # channel check of dpt.py:198-201
        f(PredictionHeadLayeredInput(list_features=feats[::-1], target_output_shape=(96, 80)))


@pytest.mark.parametrize("cls_name", ["MultiViewGlobalAttentionTransformer", "MultiViewAlternatingAttentionTransformer"])
def test_self_attention_info_sharing_at_size_vs_oracle(cls_name):
    """SURVEY 8 f2 AT SIZE: the reference's default transformer width (input 1024 -> dim 768, 12 heads, 12 blocks, RoPE) on two
    views of 32 x 32 tokens (a 2048-token global sequence), forward + every parameter gradient against the fp32 oracle at the
    1.0 x autocast + 1e-3 bar."""
    tag = f"{cls_name} 2 x 1024 tokens"
    torch.manual_seed(7)
    m = getattr(U, cls_name)(name="mv", input_embed_dim=1024, depth=12, dim=768, num_heads=12,
                             use_rand_idx_pe_for_non_reference_views=False, custom_positional_encoding=U.RoPE2D(freq=100.0)).to(DEV)
    g = torch.Generator().manual_seed(99)
    feats = [torch.randn(1, 1024, 32, 32, generator=g).to(DEV) for _ in range(2)]
    m.zero_grad(set_to_none=True)
    out = m(U.MultiViewTransformerInput(features=[f.clone().requires_grad_(True) for f in feats])).features
    sum(o.sum() for o in out).backward()
    ours_g = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_src = {k: v.detach() for k, v in m.state_dict().items() if k != "view_pos_table"}

    def run(sd):
        o = O.self_attention_info_sharing(sd, "", feats, 12, 12, alternating="Alternating" in cls_name, base=100.0,
                                          pe_for_non_ref=m.use_pe_for_non_reference_views)
        return list(o), sum(t.sum() for t in o)

    out32, g32, out16, g16 = _oracle_runs(run, sd_src)
    failures = []
    for v in range(2):
        assert torch.isfinite(out[v]).all()
        _check(tag, f"view{v}", O.parity(out[v], out32[v])[1], O.parity(out16[v], out32[v])[1], failures)
    names = [k for k, _ in m.named_parameters()]
    st = _grad_report(tag, ours_g, {k: g32[k] for k in names}, {k: g16[k] for k in names})
    for stage, (e_o, e_l) in st.items():
        _check(tag, f"grads[{stage}]", e_o, e_l, failures)
    assert not failures, failures


def test_diff_cross_attention_default_width_vs_oracle():
    """SURVEY 8 f4 AT the reference's default width: DifferentialMultiViewCrossAttentionTransformer(dim 768, 12 heads -> blocks with
    6 heads: 128-wide self-attention heads, 64-wide differential q / k against 128-wide v), depth 4, RoPE, two views of 16 x 16
    tokens; forward + every parameter gradient (lambdas and RMS sub-norm included) against the fp32 oracle.  The family runs on
    the un-fused attention, so the bar is the looser module-level one: 1.5 x autocast + 2e-3."""
    tag = "DiffCrossAttention 768/12"
    torch.manual_seed(11)
    m = U.DifferentialMultiViewCrossAttentionTransformer(name="mvd", input_embed_dim=1024, num_views=2, depth=4, dim=768, num_heads=12,
                                                         custom_positional_encoding=U.RoPE2D(freq=100.0)).to(DEV)
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(1, 1024, 16, 16, generator=g).to(DEV) for _ in range(2)]
    m.zero_grad(set_to_none=True)
    out = m(U.MultiViewTransformerInput(features=[f.clone().requires_grad_(True) for f in feats])).features
    sum(o.sum() for o in out).backward()
    ours_g = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    sd_src = {k: v.detach() for k, v in m.state_dict().items()}

    def run(sd):
        o = O.diff_info_sharing(sd, "", feats, 4, 12, base=100.0)
        return list(o), sum(t.sum() for t in o)

    out32, g32, out16, g16 = _oracle_runs(run, sd_src)
    for v in range(2):
        e_o, e_l = O.parity(out[v], out32[v])[1], O.parity(out16[v], out32[v])[1]
        print(f"{tag} view{v}: ours {e_o:.3e} vs autocast-reference {e_l:.3e}")
        assert e_o <= 1.5 * e_l + 2e-3, (v, e_o, e_l)
    names = [k for k, _ in m.named_parameters()]
    st = _grad_report(tag, ours_g, {k: g32[k] for k in names}, {k: g16[k] for k in names})
    for stage, (e_o, e_l) in st.items():
        assert e_o <= 1.5 * e_l + 2e-3, (stage, e_o, e_l)
