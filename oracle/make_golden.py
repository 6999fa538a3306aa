"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*.npz` from the REAL reference.

Run once in the build container (where /root/reference exists):
    python oracle/make_golden.py
It imports castacks/UniCeption @ 802ebc17 (pure PyTorch, CPU fp32), builds small instances of
the hot-path modules with *seeded, machine-independent* weights
(`dust3r_oracle.seeded_state_dict`, numpy RandomState), runs the reference's own forward /
backward, and stores inputs + outputs (+ gradient digests).  It also asserts that the
restatement in `oracle/dust3r_oracle.py` reproduces every vector (fp32, <= 2e-5 rel), which is
what pins the oracle.  The fixtures carry only tensors and the config -- weights are
re-derived from the seed by the tests.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dust3r_oracle as O  # noqa: E402
import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _img(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).clamp_(-1, 1)


def _load_seeded(module: nn.Module, seed: int):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = O.seeded_state_dict(shapes, seed)
    module.load_state_dict(sd)
    return {k: v.clone() for k, v in module.state_dict().items()}, shapes


def _save(name, cfg, arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(cfg), **arrays)
    print(f"wrote {name}.npz ({sum(a.nbytes for a in arrays.values())/1e6:.2f} MB raw)")


def _check(tag, got, ref, tol=2e-5):
    ma, rel = O.parity(got, ref)
    print(f"  oracle vs reference [{tag}]: max-abs {ma:.3e} rel-L2 {rel:.3e}")
    assert rel <= tol, (tag, ma, rel)


def _grad_digest(params: dict):
    """Per-parameter (sum, l2) digest of .grad, key-sorted, plus full grads of a few small tensors."""
    keys = sorted(params)
    dig = np.array([[float(params[k].grad.double().sum()), float(params[k].grad.double().norm())] for k in keys])
    return keys, dig


def golden_rope():
    from uniception.models.libs.croco.pos_embed import RoPE2D

    g = torch.Generator().manual_seed(7)
    B, H, h, w, D = 2, 3, 5, 7, 64
    tok = torch.randn(B, H, h * w, D, generator=g)
    pos = O.patch_positions(B, h, w)
    rope = RoPE2D(freq=100.0)
    out = rope(tok, pos)
    _check("rope2d fwd", O.rope2d(tok, pos, 100.0, 1.0), out, 1e-6)
    # backward of the native module == forward with -F0 (curope2d.py:24-28): check via autograd on the fallback
    t2 = tok.clone().requires_grad_(True)
    gout = torch.randn(B, H, h * w, D, generator=g)
    rope(t2, pos).backward(gout)
    _check("rope2d bwd", O.rope2d(gout, pos, 100.0, -1.0), t2.grad, 1e-6)
    arrays = dict(tokens=tok, positions=pos, out=out, grad_out=gout, grad_in=t2.grad)
    # native CPU path of the reference (curope.cpp:11-47), when oracle/_ref was built
    try:
        sys.path.insert(0, os.path.join(HERE, "_ref"))
        import curope  # type: ignore

        t3 = tok.transpose(1, 2).contiguous().clone()  # [B,N,H,D] as curope2d.py:38 passes it
        curope.rope_2d(t3, pos, 100.0, 1.0)
        _check("rope2d vs native curope CPU", O.rope2d(tok, pos, 100.0, 1.0), t3.transpose(1, 2), 1e-6)
        arrays["out_native_cpu"] = t3.transpose(1, 2).contiguous()
    except Exception as e:  # pragma: no cover
        print("  (native curope CPU not available:", repr(e)[:80], ")")
    _save("rope2d", dict(base=100.0, B=B, H=H, h=h, w=w, D=D), arrays)


def golden_index_ops():
    from uniception.models.utils.intermediate_feature_return import feature_take_indices
    from uniception.models.libs.croco.patch_embed import PositionGetter
    from uniception.models.factory.dust3r import interleave, is_symmetrized
    import torch.nn.functional as F

    pos = PositionGetter()(2, 3, 5, "cpu")
    assert torch.equal(pos, O.patch_positions(2, 3, 5))
    x = torch.arange(2 * 16 * 3 * 2, dtype=torch.float32).view(2, 16, 3, 2)
    ps = F.pixel_shuffle(x, 2)
    assert torch.equal(ps, O.pixel_shuffle(x, 2))
    cases = [(12, None), (12, 4), (12, [5, 8]), (24, [5, 11, 17, 23]), (12, [-1, -3])]
    take = []
    for n, ind in cases:
        r = feature_take_indices(n, ind)
        assert (list(r[0]), r[1]) == (O.feature_take_indices(n, ind)[0], O.feature_take_indices(n, ind)[1])
        take.append(list(r[0]))
    a, b = torch.arange(6.0).view(3, 2), -torch.arange(6.0).view(3, 2)
    i1, i2 = interleave(a, b)
    o1, o2 = O.interleave(a, b)
    assert torch.equal(i1, o1) and torch.equal(i2, o2)
    sym_cases = [([1], [2]), ([1, 2], [2, 1]), ([1, 2, 3, 4], [2, 1, 4, 3]), ([1, 2, 3, 4], [5, 6, 7, 8])]
    sym = []
    for s1, s2 in sym_cases:
        r = is_symmetrized({"instance": s1}, {"instance": s2})
        assert r == O.is_symmetrized(s1, s2)
        sym.append(bool(r))
    _save("index_ops", dict(take_cases=[[n, ind] for n, ind in cases], take=take, sym_cases=sym_cases, sym=sym),
          dict(positions_2_3_5=pos, pixel_shuffle_in=x, pixel_shuffle_out=ps, inter_a=a, inter_b=b, inter_1=i1, inter_2=i2))


def golden_encoder(name, C, depth, heads, hw, B, seed, indices=None):
    from uniception.models.encoders.base import ViTEncoderInput
    from uniception.models.encoders.croco import CroCoEncoder, CroCoIntermediateFeatureReturner

    kw = dict(name="enc", data_norm_type="dust3r", img_size=hw, enc_embed_dim=C, enc_depth=depth, enc_num_heads=heads)
    enc = CroCoEncoder(**kw) if indices is None else CroCoIntermediateFeatureReturner(indices=indices, intermediates_only=False, **kw)
    sd, shapes = _load_seeded(enc, seed)
    img = _img((B, 3, *hw), seed + 1)
    out = enc(ViTEncoderInput(image=img, data_norm_type="dust3r"))
    sdp = {"encoder." + k: v for k, v in sd.items()}
    arrays = dict(img=img)
    if indices is None:
        feat = out.features
        _check(name, O.croco_encoder(sdp, "encoder.", img, depth, heads), feat)
    else:
        feat, inter = out[0].features, [o.features for o in out[1]]
        of, oi = O.croco_encoder(sdp, "encoder.", img, depth, heads, indices=indices)
        _check(name, of, feat)
        for i, (a, b) in enumerate(zip(oi, inter)):
            _check(f"{name} inter{i}", a, b)
            arrays[f"inter{i}"] = b
    arrays["features"] = feat
    _save(name, dict(C=C, depth=depth, heads=heads, hw=list(hw), B=B, seed=seed, indices=indices,
                     shapes={k: list(v) for k, v in shapes.items()}), arrays)


def _tiny_dust3r(head, C_enc, enc_depth, enc_heads, C_dec, dec_depth, dec_heads, hw, ifr_indices=(0, 1)):
    """A real reference `DUSt3R` object with small sub-modules: bypass __init__ (which hard-codes
    ViT-L) and run the reference's own unmodified `forward` (factory/dust3r.py:250-332)."""
    from uniception.models.encoders.croco import CroCoEncoder
    from uniception.models.factory.dust3r import DUSt3R
    from uniception.models.info_sharing.cross_attention_transformer import (
        MultiViewCrossAttentionTransformer, MultiViewCrossAttentionTransformerIFR)
    from uniception.models.libs.croco.pos_embed import RoPE2D
    from uniception.models.prediction_heads.adaptors import PointMapWithConfidenceAdaptor
    from uniception.models.prediction_heads.dpt import DPTFeature, DPTRegressionProcessor
    from uniception.models.prediction_heads.linear import LinearFeature

    m = DUSt3R.__new__(DUSt3R)
    nn.Module.__init__(m)
    m.pred_head_type = head
    m.rope = RoPE2D(freq=100.0)
    m.encoder = CroCoEncoder(name="e", data_norm_type="dust3r", img_size=hw, enc_embed_dim=C_enc,
                             enc_depth=enc_depth, enc_num_heads=enc_heads)
    common = dict(name="i", input_embed_dim=C_enc, num_views=2, depth=dec_depth, dim=C_dec, num_heads=dec_heads,
                  custom_positional_encoding=m.rope)
    if head == "linear":
        m.info_sharing = MultiViewCrossAttentionTransformer(**common)
        m.head1 = LinearFeature(input_feature_dim=C_dec, output_dim=4, patch_size=16)
        m.head2 = LinearFeature(input_feature_dim=C_dec, output_dim=4, patch_size=16)
    else:
        m.info_sharing = MultiViewCrossAttentionTransformerIFR(indices=list(ifr_indices), norm_intermediate=False, **common)
        for k in (1, 2):
            f = DPTFeature(patch_size=16, hooks=[0, 1, 2, 3], input_feature_dims=[C_enc] + [C_dec] * 3,
                           layer_dims=[12, 24, 48, 96], feature_dim=32)
            r = DPTRegressionProcessor(input_feature_dim=32, output_dim=4)
            setattr(m, f"dpt_feature_head{k}", f)
            setattr(m, f"dpt_regressor_head{k}", r)
            setattr(m, f"head{k}", nn.Sequential(f, r))
    m.adaptor = PointMapWithConfidenceAdaptor(name="pointmap", pointmap_mode="exp", pointmap_vmin=-float("inf"),
                                              pointmap_vmax=float("inf"), confidence_type="exp",
                                              confidence_vmin=1, confidence_vmax=float("inf"))
    return m


def golden_dust3r(name, head, seed, B=2, hw=(32, 48), C_enc=192, enc_depth=2, enc_heads=3, C_dec=128, dec_depth=3,
                  dec_heads=2, symmetrized=False):
    m = _tiny_dust3r(head, C_enc, enc_depth, enc_heads, C_dec, dec_depth, dec_heads, hw, ifr_indices=(0, 1))
    sd, shapes = _load_seeded(m, seed)
    img1, img2 = _img((B, 3, *hw), seed + 1), _img((B, 3, *hw), seed + 2)
    if symmetrized:  # pairs (a,b),(b,a)
        img1 = torch.stack((img1[0], img2[0]), 0).repeat(B // 2, 1, 1, 1)
        img2 = torch.stack((img2[0], img1[0]), 0).repeat(B // 2, 1, 1, 1)
        inst1, inst2 = ["a", "b"] * (B // 2), ["b", "a"] * (B // 2)
    else:
        inst1, inst2 = [str(i) for i in range(B)], [str(B + i) for i in range(B)]
    v1 = {"img": img1, "instance": inst1, "data_norm_type": "dust3r"}
    v2 = {"img": img2, "instance": inst2, "data_norm_type": "dust3r"}
    r1, r2 = m(v1, v2)
    loss = O.bench_loss(r1, r2)
    loss.backward()
    params = dict(m.named_parameters())  # (aliases de-duplicated by named_parameters)
    gkeys, gdig = _grad_digest(params)

    # the oracle on the same weights / inputs
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    if head == "dpt":  # oracle heads read the un-aliased names
        pass
    o1, o2 = O.dust3r_forward(osd, img1, img2, enc_depth=enc_depth, enc_heads=enc_heads, dec_depth=dec_depth,
                              dec_heads=dec_heads, head=head, instances=(inst1, inst2)) if head == "linear" else \
        _oracle_dpt_forward(osd, img1, img2, enc_depth, enc_heads, dec_depth, dec_heads)
    for k in r1:
        _check(f"{name} res1.{k}", o1[k], r1[k])
    for k in r2:
        _check(f"{name} res2.{k}", o2[k], r2[k])
    O.bench_loss(o1, o2).backward()
    for k in ("encoder.patch_embed.proj.weight", "encoder.enc_blocks.0.attn.qkv.weight",
              "info_sharing.multi_view_branches.1.0.cross_attn.projk.weight", "info_sharing.norm.weight"):
        _check(f"{name} grad {k}", osd[k].grad, params[k].grad, 1e-4)
    arrays = dict(img1=img1, img2=img2, pts3d_1=r1["pts3d"], conf_1=r1["conf"], pts3d_2=r2["pts3d_in_other_view"],
                  conf_2=r2["conf"], loss=loss.detach(), grad_digest=gdig,
                  grad_qkv0=params["encoder.enc_blocks.0.attn.qkv.weight"].grad,
                  grad_patch=params["encoder.patch_embed.proj.weight"].grad,
                  grad_projk=params["info_sharing.multi_view_branches.1.0.cross_attn.projk.weight"].grad)
    _save(name, dict(head=head, seed=seed, B=B, hw=list(hw), C_enc=C_enc, enc_depth=enc_depth, enc_heads=enc_heads,
                     C_dec=C_dec, dec_depth=dec_depth, dec_heads=dec_heads, inst1=inst1, inst2=inst2,
                     ifr_indices=[0, 1], grad_keys=gkeys, shapes={k: list(v) for k, v in shapes.items()}), arrays)


def _oracle_dpt_forward(sd, img1, img2, enc_depth, enc_heads, dec_depth, dec_heads):
    """dust3r_forward(head='dpt') but with the tiny fixture's IFR indices (0,1) instead of (5,8)."""
    B, _, H, W = img1.shape
    feat = O.croco_encoder(sd, "encoder.", torch.cat((img1, img2), 0), enc_depth, enc_heads)
    f1, f2 = feat.chunk(2, dim=0)
    (d1, d2), inter = O.info_sharing(sd, "info_sharing.", [f1, f2], dec_depth, dec_heads, indices=[0, 1],
                                     norm_intermediate=False)
    o1 = O.dpt_regressor(sd, "dpt_regressor_head1.", O.dpt_feature(sd, "dpt_feature_head1.", [f1, inter[0][0], inter[1][0], d1]), (H, W))
    o2 = O.dpt_regressor(sd, "dpt_regressor_head2.", O.dpt_feature(sd, "dpt_feature_head2.", [f2, inter[0][1], inter[1][1], d2]), (H, W))
    p1, c1 = O.pointmap_conf_adaptor(o1)
    p2, c2 = O.pointmap_conf_adaptor(o2)
    return ({"pts3d": p1.permute(0, 2, 3, 1), "conf": c1.permute(0, 2, 3, 1)},
            {"pts3d_in_other_view": p2.permute(0, 2, 3, 1), "conf": c2.permute(0, 2, 3, 1)})


def golden_depth_c5(name, seed, B=2, hw=(42, 56), patch=14, C=128, depth=4, heads=2, indices=(0, 1, 2, 3)):
    """BASELINE.json configs[4] in miniature: ViT encoder with PATCH 14 (intermediate-feature returner) -> DPTFeature ->
    DPTRegressionProcessor -> DepthAdaptor(exp), all the reference's own modules (the in-tree CroCo encoder
    stands in for the un-vendored DINOv2; prediction_heads/dpt.py:180-311, adaptors.py:233-257)."""
    from uniception.models.encoders.base import ViTEncoderInput
    from uniception.models.encoders.croco import CroCoIntermediateFeatureReturner
    from uniception.models.prediction_heads.adaptors import DepthAdaptor
    from uniception.models.prediction_heads.base import AdaptorInput, PredictionHeadLayeredInput
    from uniception.models.prediction_heads.dpt import DPTFeature, DPTRegressionProcessor

    m = nn.Module()
    m.encoder = CroCoIntermediateFeatureReturner(name="enc", data_norm_type="dust3r", img_size=hw, patch_size=patch, enc_embed_dim=C,
                                                 enc_depth=depth, enc_num_heads=heads, indices=list(indices), intermediates_only=True)
    m.dpt_feature_head = DPTFeature(patch_size=patch, hooks=[0, 1, 2, 3], input_feature_dims=[C] * 4, layer_dims=[12, 24, 48, 96],
                                    feature_dim=32)
    m.dpt_regressor_head = DPTRegressionProcessor(input_feature_dim=32, output_dim=1)
    adaptor = DepthAdaptor(name="depth", mode="exp")
    sd, shapes = _load_seeded(m, seed)
    img = _img((B, 3, *hw), seed + 1)
    feats = [o.features for o in m.encoder(ViTEncoderInput(image=img, data_norm_type="dust3r"))]
    dense = m.dpt_feature_head(PredictionHeadLayeredInput(list_features=feats, target_output_shape=hw))
    raw = m.dpt_regressor_head(dense).decoded_channels
    out = adaptor(AdaptorInput(adaptor_feature=raw, output_shape_hw=hw)).value
    out.sum().backward()
    params = dict(m.named_parameters())
    # oracle on the same (alias-resolved) weights
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    _, ointer = O.croco_encoder(osd, "encoder.", img, depth, heads, patch, indices=list(indices), norm_intermediate=True)
    oraw = O.dpt_regressor(osd, "dpt_regressor_head.", O.dpt_feature(osd, "dpt_feature_head.", ointer), hw)
    oout = O.depth_adaptor(oraw, "exp")
    for i, (a, b) in enumerate(zip(ointer, feats)):
        _check(f"{name} hook{i}", a, b)
    _check(f"{name} raw", oraw, raw)
    _check(f"{name} depth", oout, out)
    oout.sum().backward()
    k0 = "encoder.enc_blocks.0.attn.qkv.weight"
    _check(f"{name} grad {k0}", osd[k0].grad, params[k0].grad, 1e-4)
    _check(f"{name} grad patch_embed", osd["encoder.patch_embed.proj.weight"].grad, params["encoder.patch_embed.proj.weight"].grad, 1e-4)
    _save(name, dict(seed=seed, B=B, hw=list(hw), patch=patch, C=C, depth=depth, heads=heads, indices=list(indices),
                     shapes={k: list(v) for k, v in shapes.items()}),
          dict(img=img, raw=raw, depth=out, hook0=feats[0], hook3=feats[3], grad_qkv0=params[k0].grad,
               grad_patch=params["encoder.patch_embed.proj.weight"].grad))


def golden_encoder_manyar(name, seed, C=128, depth=2, heads=2, hw=(32, 48), B=3):
    """Mixed aspect-ratio batch through the reference's `ManyAR_PatchEmbed` encoder (patch_embed.py:85-127, croco.py:160-168):
    samples 0 and 2 landscape, sample 1 portrait (stored transposed, true_shape = (W, H))."""
    from uniception.models.encoders.base import ViTEncoderInput
    from uniception.models.encoders.croco import CroCoEncoder

    enc = CroCoEncoder(name="enc", data_norm_type="dust3r", patch_embed_cls="ManyAR_PatchEmbed", img_size=hw, enc_embed_dim=C,
                       enc_depth=depth, enc_num_heads=heads)
    sd, shapes = _load_seeded(enc, seed)
    img = _img((B, 3, *hw), seed + 1)
    true_shape = torch.tensor([[hw[0], hw[1]], [hw[1], hw[0]], [hw[0], hw[1]]][:B])
    inp = ViTEncoderInput(image=img, data_norm_type="dust3r")
    inp.true_shape = true_shape
    feat = enc(inp).features
    feat.sum().backward()
    params = dict(enc.named_parameters())
    sdp = {"encoder." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    of = O.croco_encoder(sdp, "encoder.", img, depth, heads, true_shape=true_shape)
    _check(name, of, feat)
    of.sum().backward()
    _check(name + " grad patch", sdp["encoder.patch_embed.proj.weight"].grad, params["patch_embed.proj.weight"].grad, 1e-4)
    _check(name + " grad qkv0", sdp["encoder.enc_blocks.0.attn.qkv.weight"].grad, params["enc_blocks.0.attn.qkv.weight"].grad, 1e-4)
    _save(name, dict(C=C, depth=depth, heads=heads, hw=list(hw), B=B, seed=seed, shapes={k: list(v) for k, v in shapes.items()}),
          dict(img=img, true_shape=true_shape, features=feat, grad_patch=params["patch_embed.proj.weight"].grad,
               grad_qkv0=params["enc_blocks.0.attn.qkv.weight"].grad))


def golden_self_attention_info_sharing(name, cls_name, seed, rope, V=2, B=2, hw=(3, 4), C_in=192, dim=128, depth=4, heads=2,
                                       scaling=False, indices=None, norm_intermediate=True, qk_norm=False, init_values=None):
    """`MultiViewGlobalAttentionTransformer` / `MultiViewAlternatingAttentionTransformer` on V views, with
    sequential view-index positional encodings (the default draws them at random) and optional RoPE.  With `indices` the
    class is the `...IFR` variant and the loss weights intermediate level k by (0.5 + k)."""
    from uniception.models.info_sharing import alternating_attention_transformer as AT, global_attention_transformer as GT
    from uniception.models.info_sharing.base import MultiViewTransformerInput

    from uniception.models.libs.croco.pos_embed import RoPE2D

    cls = getattr(GT, cls_name, None) or getattr(AT, cls_name)
    # a callable, not the string "rope": the alternating transformer does not resolve the string (it would call a str)
    m = cls(name="mv", input_embed_dim=C_in, depth=depth, dim=dim, num_heads=heads, use_rand_idx_pe_for_non_reference_views=False,
            custom_positional_encoding=RoPE2D(freq=100.0) if rope else None,
            use_scalable_softmax=scaling, use_entropy_scaling=scaling, qk_norm=qk_norm, init_values=init_values,
            **(dict(indices=list(indices), norm_intermediate=norm_intermediate) if indices is not None else {}))
    sm = (True, True, m.base_token_count_for_entropy_scaling, m.entropy_scaling_growth_factor) if scaling else None
    # seeded parameters only: `view_pos_table` is a persistent BUFFER (the sinusoid table), not a weight
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k != "view_pos_table"}
    m.load_state_dict(O.seeded_state_dict(shapes, seed), strict=False)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    feats = [torch.randn(B, C_in, *hw, generator=g).requires_grad_(True) for _ in range(V)]
    res = m(MultiViewTransformerInput(features=feats))
    inter = []
    if indices is not None:
        res, inter = res
        inter = [lvl.features for lvl in inter]
    out = res.features
    (sum(o.sum() for o in out) + sum((0.5 + k) * sum(t.sum() for t in lvl) for k, lvl in enumerate(inter))).backward()
    params = dict(m.named_parameters())
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "view_pos_table"}
    of = [f.detach().clone().requires_grad_(True) for f in feats]
    oo = O.self_attention_info_sharing(osd, "", of, depth, heads, alternating="Alternating" in cls_name, base=100.0 if rope else None,
                                       distinguish_ref=m.distinguish_ref_and_non_ref_views, pe_for_non_ref=m.use_pe_for_non_reference_views,
                                       softmax_scaling=sm, indices=indices, norm_intermediate=norm_intermediate)
    oi = []
    if indices is not None:
        oo, oi = oo
    for v in range(V):
        _check(f"{name} view{v}", oo[v], out[v])
    for k, lvl in enumerate(oi):
        for v in range(V):
            _check(f"{name} inter{k} view{v}", lvl[v], inter[k][v])
    (sum(o.sum() for o in oo) + sum((0.5 + k) * sum(t.sum() for t in lvl) for k, lvl in enumerate(oi))).backward()
    k0 = "self_attention_blocks.1.attn.qkv.weight"
    _check(f"{name} grad {k0}", osd[k0].grad, params[k0].grad, 1e-4)
    _check(f"{name} grad input0", of[0].grad, feats[0].grad, 1e-4)
    arrays = {f"feat{v}": feats[v].detach() for v in range(V)}
    arrays.update({f"out{v}": out[v] for v in range(V)})
    arrays.update({f"inter{k}_{v}": inter[k][v] for k in range(len(inter)) for v in range(V)})
    arrays.update(grad_qkv1=params[k0].grad, grad_proj_embed=params["proj_embed.weight"].grad, grad_in0=feats[0].grad,
                  grad_norm_w=params["norm.weight"].grad, grad_fc2_b1=params["self_attention_blocks.1.mlp.fc2.bias"].grad)
    for extra in ("attn.q_norm.weight", "attn.k_norm.bias", "ls1.gamma", "ls2.gamma", "attn.proj.bias"):
        kx = "self_attention_blocks.1." + extra
        if kx in params and (qk_norm or init_values):
            _check(f"{name} grad {kx}", osd[kx].grad, params[kx].grad, 1e-4)
            arrays["grad_" + extra.replace(".", "_")] = params[kx].grad
    _save(name, dict(cls=cls_name, seed=seed, rope=bool(rope), V=V, B=B, hw=list(hw), C_in=C_in, dim=dim, depth=depth, heads=heads,
                     pe_for_non_ref=bool(m.use_pe_for_non_reference_views), scaling=bool(scaling),
                     indices=list(indices) if indices is not None else None, norm_intermediate=bool(norm_intermediate),
                     qk_norm=bool(qk_norm), init_values=init_values,
                     shapes={k: list(v) for k, v in shapes.items()}), arrays)


def golden_self_attention_block(name, seed, B=2, hw=(4, 5), dim=192, latent=128, heads=2):
    """A stand-alone `SelfAttentionBlock` (utils/transformer_blocks.py:415-514) with every built option at once:
    latent_attn_dim, qk_norm, LayerScale, RoPE."""
    from functools import partial

    from uniception.models.libs.croco.pos_embed import RoPE2D
    from uniception.models.utils.transformer_blocks import SelfAttentionBlock

    m = SelfAttentionBlock(dim=dim, num_heads=heads, latent_attn_dim=latent, qkv_bias=True, qk_norm=True, init_values=0.5,
                           norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), custom_positional_encoding=RoPE2D(freq=100.0))
    sd, shapes = _load_seeded(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, hw[0] * hw[1], dim, generator=g).requires_grad_(True)
    pos = O.patch_positions(B, hw[0], hw[1], "cpu")
    y = m(x, pos)
    y.sum().backward()
    params = dict(m.named_parameters())
    osd = {"b." + k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ox = x.detach().clone().requires_grad_(True)
    oy = O.encoder_block(osd, "b.", ox, pos, heads, 100.0)
    _check(f"{name} out", oy, y)
    oy.sum().backward()
    arrays = dict(x=x.detach(), y=y, grad_x=x.grad)
    for k in ("attn.qkv.weight", "attn.q_norm.weight", "attn.proj.weight", "ls1.gamma", "ls2.gamma", "mlp.fc1.bias"):
        _check(f"{name} grad {k}", osd["b." + k].grad, params[k].grad, 1e-4)
        arrays["grad_" + k.replace(".", "_")] = params[k].grad
    _check(f"{name} grad x", ox.grad, x.grad, 1e-4)
    _save(name, dict(seed=seed, B=B, hw=list(hw), dim=dim, latent=latent, heads=heads, shapes={k: list(v) for k, v in shapes.items()}), arrays)


def golden_additional_tokens(name, cls_name, seed, V=2, T=2, Tv=1, B=2, hw=(3, 4), C_in=192, dim=128, depth=4, heads=2, indices=None):
    """Additional input tokens in the global / alternating attention transformers (global_attention_transformer.py:266-333,
    :434-461; alternating_attention_transformer.py:402-447): T global tokens and Tv tokens per view, no positional encoding
    plugin (the reference refuses RoPE with additional tokens).  The loss weights maps 1, global extras 2, per-view extras 3
    (and intermediate level k by 0.5 + k on top)."""
    from uniception.models.info_sharing import alternating_attention_transformer as AT, global_attention_transformer as GT
    from uniception.models.info_sharing.base import MultiViewTransformerInput

    cls = getattr(GT, cls_name, None) or getattr(AT, cls_name)
    m = cls(name="mv", input_embed_dim=C_in, depth=depth, dim=dim, num_heads=heads, use_rand_idx_pe_for_non_reference_views=False,
            **(dict(indices=list(indices)) if indices is not None else {}))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k != "view_pos_table"}
    m.load_state_dict(O.seeded_state_dict(shapes, seed), strict=False)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    feats = [torch.randn(B, C_in, *hw, generator=g).requires_grad_(True) for _ in range(V)]
    extra = torch.randn(B, C_in, T, generator=g).requires_grad_(True) if T else None
    per_view = [torch.randn(B, C_in, Tv, generator=g).requires_grad_(True) for _ in range(V)] if Tv else None

    def level_loss(maps, ex, pv):
        return sum(t.sum() for t in maps) + (2 * ex.sum() if ex is not None else 0) + (3 * sum(t.sum() for t in pv) if pv else 0)

    def unpack(o):
        return o.features, o.additional_token_features, o.additional_token_features_per_view

    res = m(MultiViewTransformerInput(features=feats, additional_input_tokens=extra, additional_input_tokens_per_view=per_view))
    inter = []
    if indices is not None:
        res, inter = res
    levels = [unpack(res)] + [unpack(o) for o in inter]
    sum((1.0 if k == 0 else k - 0.5) * level_loss(*lv) for k, lv in enumerate(levels)).backward()
    params = dict(m.named_parameters())
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "view_pos_table"}
    of = [f.detach().clone().requires_grad_(True) for f in feats]
    oe = extra.detach().clone().requires_grad_(True) if T else None
    opv = [t.detach().clone().requires_grad_(True) for t in per_view] if Tv else None
    oo = O.self_attention_info_sharing(osd, "", of, depth, heads, alternating="Alternating" in cls_name,
                                       distinguish_ref=m.distinguish_ref_and_non_ref_views, pe_for_non_ref=m.use_pe_for_non_reference_views,
                                       indices=indices, extra=oe, extra_per_view=opv)
    olevels = [oo] if indices is None else [oo[0]] + list(oo[1])
    arrays = {f"feat{v}": feats[v].detach() for v in range(V)}
    for k, (lv, olv) in enumerate(zip(levels, olevels)):
        for v in range(V):
            _check(f"{name} level{k} view{v}", olv[0][v], lv[0][v])
            arrays[f"l{k}_out{v}"] = lv[0][v]
        if T:
            _check(f"{name} level{k} global extras", olv[1], lv[1])
            arrays[f"l{k}_extra"] = lv[1]
        if Tv:
            for v in range(V):
                _check(f"{name} level{k} per-view extras {v}", olv[2][v], lv[2][v])
                arrays[f"l{k}_pv{v}"] = lv[2][v]
    sum((1.0 if k == 0 else k - 0.5) * level_loss(*lv) for k, lv in enumerate(olevels)).backward()
    k0 = "self_attention_blocks.1.attn.qkv.weight"
    _check(f"{name} grad {k0}", osd[k0].grad, params[k0].grad, 1e-4)
    _check(f"{name} grad input0", of[0].grad, feats[0].grad, 1e-4)
    arrays.update(grad_qkv1=params[k0].grad, grad_proj_embed=params["proj_embed.weight"].grad, grad_in0=feats[0].grad,
                  grad_fc2_b1=params["self_attention_blocks.1.mlp.fc2.bias"].grad, grad_fc2_b0=params["self_attention_blocks.0.mlp.fc2.bias"].grad,
                  grad_proj_embed_b=params["proj_embed.bias"].grad)
    if T:
        _check(f"{name} grad global extras", oe.grad, extra.grad, 1e-4)
        arrays.update(extra=extra.detach(), grad_extra=extra.grad)
    if Tv:
        _check(f"{name} grad per-view extras", opv[V - 1].grad, per_view[V - 1].grad, 1e-4)
        arrays.update({f"pv{v}": per_view[v].detach() for v in range(V)})
        arrays["grad_pv_last"] = per_view[V - 1].grad
    _save(name, dict(cls=cls_name, seed=seed, V=V, T=T, Tv=Tv, B=B, hw=list(hw), C_in=C_in, dim=dim, depth=depth, heads=heads,
                     pe_for_non_ref=bool(m.use_pe_for_non_reference_views), indices=list(indices) if indices is not None else None,
                     shapes={k: list(v) for k, v in shapes.items()}), arrays)


def golden_cross_attention_scaled(name, seed, B=2, hw=(3, 4), C_in=192, dim=128, depth=2, heads=2, scaling=True, qk_norm=False,
                                  init_values=None):
    """`MultiViewCrossAttentionTransformer(use_scalable_softmax=True, use_entropy_scaling=True)` (the two
    token-count-dependent query multipliers, utils/transformer_blocks.py:231-241, :360-370); with `qk_norm` / `init_values`
    the per-head q/k LayerNorms (:199-200, :306-307) and LayerScale (:389-412) instead."""
    from uniception.models.info_sharing.base import MultiViewTransformerInput
    from uniception.models.info_sharing.cross_attention_transformer import MultiViewCrossAttentionTransformer
    from uniception.models.libs.croco.pos_embed import RoPE2D

    m = MultiViewCrossAttentionTransformer(name="mv", input_embed_dim=C_in, num_views=2, depth=depth, dim=dim, num_heads=heads,
                                           custom_positional_encoding=RoPE2D(freq=100.0), use_scalable_softmax=scaling,
                                           use_entropy_scaling=scaling, qk_norm=qk_norm, init_values=init_values)
    sd, shapes = _load_seeded(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    feats = [torch.randn(B, C_in, *hw, generator=g).requires_grad_(True) for _ in range(2)]
    out = m(MultiViewTransformerInput(features=feats)).features
    sum(o.sum() for o in out).backward()
    params = dict(m.named_parameters())
    sm = (True, True, m.base_token_count_for_entropy_scaling, m.entropy_scaling_growth_factor) if scaling else None
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    of = [f.detach().clone().requires_grad_(True) for f in feats]
    oo = O.info_sharing(osd, "", of, depth, heads, softmax_scaling=sm)
    for v in range(2):
        _check(f"{name} view{v}", oo[v], out[v])
    sum(o.sum() for o in oo).backward()
    k0 = "multi_view_branches.1.0.cross_attn.projq.weight"
    _check(f"{name} grad {k0}", osd[k0].grad, params[k0].grad, 1e-4)
    arrays = {f"feat{v}": feats[v].detach() for v in range(2)}
    arrays.update({f"out{v}": out[v] for v in range(2)})
    arrays.update(grad_projq=params[k0].grad, grad_qkv=params["multi_view_branches.0.1.attn.qkv.weight"].grad, grad_in0=feats[0].grad)
    for extra in ("cross_attn.q_norm.weight", "cross_attn.k_norm.bias", "attn.q_norm.bias", "ls1.gamma", "ls2.gamma", "ls3.gamma",
                  "cross_attn.proj.bias", "mlp.fc2.bias"):
        kx = "multi_view_branches.1.0." + extra
        if kx in params and (qk_norm or init_values):
            _check(f"{name} grad {kx}", osd[kx].grad, params[kx].grad, 1e-4)
            arrays["grad_" + extra.replace(".", "_")] = params[kx].grad
    _save(name, dict(seed=seed, B=B, hw=list(hw), C_in=C_in, dim=dim, depth=depth, heads=heads, softmax_scaling=list(sm) if sm else None,
                     qk_norm=bool(qk_norm), init_values=init_values,
                     shapes={k: list(v) for k, v in shapes.items()}), arrays)


def golden_diff_cross_attention(name, seed, B=2, hw=(3, 4), C_in=192, dim=256, depth=2, heads=4, V=2, rope=True, indices=None):
    """`DifferentialMultiViewCrossAttentionTransformer(IFR)` (diff_cross_attention_transformer.py:22-588):
    dim 256 / 4 heads -> blocks with 2 heads: 128-wide self-attention heads, differential cross-attention with 64-wide q / k
    against 128-wide v (the head geometry of the default dim 768 / 12 heads).  Loss weights every output."""
    from uniception.models.info_sharing.base import MultiViewTransformerInput
    from uniception.models.info_sharing.diff_cross_attention_transformer import (
        DifferentialMultiViewCrossAttentionTransformer, DifferentialMultiViewCrossAttentionTransformerIFR)
    from uniception.models.libs.croco.pos_embed import RoPE2D

    kw = dict(name="mvd", input_embed_dim=C_in, num_views=V, depth=depth, dim=dim, num_heads=heads,
              custom_positional_encoding=RoPE2D(freq=100.0) if rope else None)
    if indices is None:
        m = DifferentialMultiViewCrossAttentionTransformer(**kw)
    else:
        m = DifferentialMultiViewCrossAttentionTransformerIFR(indices=indices, **kw)
    sd, shapes = _load_seeded(m, seed)
    g = torch.Generator().manual_seed(seed + 1)
    feats = [torch.randn(B, C_in, *hw, generator=g).requires_grad_(True) for _ in range(V)]
    res = m(MultiViewTransformerInput(features=feats))
    out, inter = (res[0].features, [lv.features for lv in res[1]]) if indices is not None else (res.features, [])

    def loss(o, it):
        return sum(t.sum() for t in o) + sum((k + 1.5) * sum(t.sum() for t in lv) for k, lv in enumerate(it))

    loss(out, inter).backward()
    params = dict(m.named_parameters())
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    of = [f.detach().clone().requires_grad_(True) for f in feats]
    ores = O.diff_info_sharing(osd, "", of, depth, heads, base=100.0 if rope else None, indices=indices)
    oo, oi = (ores[0], ores[1]) if indices is not None else (ores, [])
    for v in range(V):
        _check(f"{name} view{v}", oo[v], out[v])
    for k, lv in enumerate(oi):
        for v in range(V):
            _check(f"{name} level{k} view{v}", lv[v], inter[k][v])
    loss(oo, oi).backward()
    arrays = {f"feat{v}": feats[v].detach() for v in range(V)}
    arrays.update({f"out{v}": out[v] for v in range(V)})
    for k, lv in enumerate(inter):
        arrays.update({f"inter{k}_{v}": lv[v] for v in range(V)})
    arrays["grad_in0"] = feats[0].grad
    for key in ("multi_view_branches.1.0.cross_attn.projq.weight", "multi_view_branches.0.1.attn.qkv.weight",
                "multi_view_branches.0.0.cross_attn.lambda_q1", "multi_view_branches.1.1.cross_attn.lambda_k2",
                "multi_view_branches.0.1.cross_attn.subln.weight", "multi_view_branches.1.0.cross_attn.projv.bias",
                "multi_view_branches.0.0.mlp.fc1.weight", "proj_embed.weight"):
        _check(f"{name} grad {key}", osd[key].grad, params[key].grad, 1e-4)
        arrays["grad_" + key.replace(".", "_")] = params[key].grad
    _save(name, dict(seed=seed, B=B, hw=list(hw), C_in=C_in, dim=dim, depth=depth, heads=heads, V=V, rope=bool(rope),
                     indices=list(indices) if indices is not None else None,
                     shapes={k: list(v) for k, v in shapes.items()}), arrays)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_import.import_reference()
    golden_index_ops()
    golden_rope()
    golden_encoder("encoder_tiny", C=128, depth=2, heads=2, hw=(32, 48), B=2, seed=11)
    golden_encoder("encoder_tiny_ifr", C=128, depth=4, heads=2, hw=(48, 32), B=1, seed=12, indices=[1, 3])
    # BASELINE.json configs[0]: ViT-B/16 encoder, one 224x224 image, CPU
    golden_encoder("encoder_vitb16_224", C=768, depth=12, heads=12, hw=(224, 224), B=1, seed=13)
    golden_dust3r("dust3r_tiny_linear", "linear", seed=21)
    golden_dust3r("dust3r_tiny_linear_sym", "linear", seed=22, B=4, symmetrized=True)
    golden_dust3r("dust3r_tiny_dpt", "dpt", seed=23, hw=(32, 32))
    golden_depth_c5("depth_c5_tiny_patch14", seed=31)
    golden_encoder_manyar("encoder_tiny_manyar", seed=41)
    golden_self_attention_info_sharing("global_attn_tiny", "MultiViewGlobalAttentionTransformer", seed=51, rope=False)
    golden_self_attention_info_sharing("global_attn_tiny_rope", "MultiViewGlobalAttentionTransformer", seed=52, rope=True, V=3)
    golden_self_attention_info_sharing("alternating_attn_tiny", "MultiViewAlternatingAttentionTransformer", seed=53, rope=True)
    golden_self_attention_info_sharing("global_attn_tiny_scaled", "MultiViewGlobalAttentionTransformer", seed=54, rope=True, scaling=True)
    golden_cross_attention_scaled("cross_attn_tiny_scaled", seed=55)
    golden_cross_attention_scaled("cross_attn_tiny_qknorm_ls", seed=58, scaling=False, qk_norm=True, init_values=0.5)
    golden_self_attention_info_sharing("alternating_attn_tiny_qknorm_ls", "MultiViewAlternatingAttentionTransformer", seed=59,
                                       rope=True, V=3, qk_norm=True, init_values=0.5)
    golden_self_attention_block("self_attn_block_latent_qknorm_ls", seed=63)
    golden_additional_tokens("global_attn_tiny_tokens", "MultiViewGlobalAttentionTransformer", seed=60, V=3, T=2, Tv=1)
    golden_additional_tokens("alternating_attn_tiny_tokens", "MultiViewAlternatingAttentionTransformerIFR", seed=61, V=2, T=3, Tv=2,
                             indices=[0, 1])
    golden_additional_tokens("alternating_attn_tiny_pv_tokens", "MultiViewAlternatingAttentionTransformer", seed=62, V=2, T=0, Tv=2)
    golden_self_attention_info_sharing("global_attn_tiny_ifr", "MultiViewGlobalAttentionTransformerIFR", seed=56, rope=True,
                                       indices=[1, 3])
    golden_self_attention_info_sharing("alternating_attn_tiny_ifr", "MultiViewAlternatingAttentionTransformerIFR", seed=57, rope=True,
                                       V=3, indices=[0, 2], norm_intermediate=False)
    golden_diff_cross_attention("diff_cross_attn_tiny", seed=71)
    golden_diff_cross_attention("diff_cross_attn_tiny_ifr", seed=72, V=3, hw=(2, 5), indices=[0, 1], rope=False)


if __name__ == "__main__":
    main()
