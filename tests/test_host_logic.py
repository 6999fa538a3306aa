"""CPU: host-side logic of the drop-in modules (construction, state-dict keys/shapes/order identical to
the reference, index ops, error behaviour).  No kernel is launched."""
import pytest
import torch

import uniception_b200 as U
from golden_utils import load


def _tiny(cfg):
    return U.DUSt3R(name="t", img_size=tuple(cfg["hw"]),
                    encoder_kwargs=dict(enc_embed_dim=cfg["C_enc"], enc_depth=cfg["enc_depth"], enc_num_heads=cfg["enc_heads"]),
                    info_sharing_kwargs=dict(depth=cfg["dec_depth"], dim=cfg["C_dec"], num_heads=cfg["dec_heads"]))


def test_state_dict_matches_reference_keys_shapes_order():
    cfg, _ = load("dust3r_tiny_linear")
    sd = _tiny(cfg).state_dict()
    assert list(sd.keys()) == list(cfg["shapes"].keys())
    for k, v in sd.items():
        assert list(v.shape) == cfg["shapes"][k], k


def test_encoder_state_dict_matches_reference():
    cfg, _ = load("encoder_vitb16_224")
    enc = U.CroCoEncoder(name="e", data_norm_type="dust3r", img_size=(224, 224), enc_embed_dim=768, enc_depth=12, enc_num_heads=12)
    sd = enc.state_dict()
    assert list(sd.keys()) == list(cfg["shapes"].keys())
    assert all(list(sd[k].shape) == cfg["shapes"][k] for k in sd)
    assert enc.name == "e" and enc.data_norm_type == "dust3r" and enc.patch_size == 16 and enc.enc_embed_dim == 768


def test_full_size_model_parameter_count():
    m = U.DUSt3R(name="dust3r", img_size=(512, 512))
    n = sum(p.numel() for p in m.parameters())
    assert n == 532_342_016 and len(m.state_dict()) == 876  # 532.34 M params, 876 tensors


def test_index_ops_match_reference_goldens():
    cfg, a = load("index_ops")
    for (n, ind), take in zip(cfg["take_cases"], cfg["take"]):
        assert U.feature_take_indices(n, ind)[0] == take
    i1, i2 = U.interleave(a["inter_a"], a["inter_b"])
    assert torch.equal(i1, a["inter_1"]) and torch.equal(i2, a["inter_2"])
    for (s1, s2), r in zip(cfg["sym_cases"], cfg["sym"]):
        assert U.is_symmetrized({"instance": s1}, {"instance": s2}) == r
    from uniception_b200.encoders import PositionGetter
    from uniception_b200.engine import grid_positions

    assert torch.equal(PositionGetter()(2, 3, 5, "cpu"), a["positions_2_3_5"])
    assert torch.equal(grid_positions(2, 3, 5, "cpu", torch.int64).view(2, 15, 2), a["positions_2_3_5"])


def test_error_behaviour_matches_reference():
    enc = U.CroCoEncoder(name="e", data_norm_type="dust3r", img_size=(32, 32), enc_embed_dim=128, enc_depth=1, enc_num_heads=2)
    with pytest.raises(AssertionError):  # encoders/base.py:94-96
        enc(U.ViTEncoderInput(image=torch.zeros(1, 3, 32, 32), data_norm_type="imagenet"))
    with pytest.raises(RuntimeError):  # CUDA only, no CPU fallback
        enc(U.ViTEncoderInput(image=torch.zeros(1, 3, 32, 32), data_norm_type="dust3r"))
    with pytest.raises(NotImplementedError):  # croco.py:86-87
        U.CroCoEncoder(name="e", data_norm_type="dust3r", pos_embed="cosine")
    with pytest.raises(ValueError):  # encoders/__init__.py:109-112
        U.encoder_factory("nope")
    dec = U.MultiViewCrossAttentionTransformer(name="i", input_embed_dim=128, num_views=2, depth=1, dim=128, num_heads=2)
    with pytest.raises(AssertionError):  # cross_attention_transformer.py:207-215
        dec(U.MultiViewTransformerInput(features=[torch.zeros(1, 128, 2, 2)]))
    with pytest.raises(AssertionError):
        dec(U.MultiViewTransformerInput(features=[torch.zeros(1, 64, 2, 2), torch.zeros(1, 64, 2, 2)]))
    with pytest.raises(ValueError):  # dust3r.py:146
        U.DUSt3R(name="x", pred_head_type="nope")
    head = U.LinearFeature(input_feature_dim=128, output_dim=4, patch_size=16)
    with pytest.raises(AssertionError):  # linear.py:77-79
        head(U.PredictionHeadInput(last_feature=torch.zeros(1, 64, 2, 2)))


def test_self_attention_transformers_host_logic():
    """Global / alternating attention transformers: registry, option containers, input validation and error
    behaviour of the reference (global_attention_transformer.py:245-351) -- all before any kernel would launch."""
    kw = dict(name="g", input_embed_dim=128, depth=2, dim=128, num_heads=2)
    assert U.INFO_SHARING_CLASSES["global_attention"] == (U.MultiViewGlobalAttentionTransformer, U.MultiViewGlobalAttentionTransformerIFR)
    assert U.INFO_SHARING_CLASSES["alternating_attention"] == (U.MultiViewAlternatingAttentionTransformer,
                                                              U.MultiViewAlternatingAttentionTransformerIFR)
    g = U.MultiViewGlobalAttentionTransformer(**kw)
    assert g.use_pe_for_non_reference_views and not U.MultiViewAlternatingAttentionTransformer(**kw).use_pe_for_non_reference_views
    assert list(g.state_dict())[0] == "view_pos_table" and g.view_pos_table.shape == (1000, 128)
    assert isinstance(g.proj_embed, torch.nn.Identity)  # input_embed_dim == dim (global_attention_transformer.py:150-153)
    f = [torch.zeros(1, 128, 2, 2), torch.zeros(1, 128, 2, 2)]
    with pytest.raises(AssertionError):  # channel mismatch
        g(U.MultiViewTransformerInput(features=[torch.zeros(1, 64, 2, 2)] * 2))
    with pytest.raises(AssertionError):  # one token tensor per view
        g(U.MultiViewTransformerInput(features=f, additional_input_tokens_per_view=[torch.zeros(1, 128, 1)]))
    with pytest.raises(AssertionError):  # (N, C, T)
        g(U.MultiViewTransformerInput(features=f, additional_input_tokens=torch.zeros(1, 128)))
    with pytest.raises(AssertionError):  # batch mismatch
        g(U.MultiViewTransformerInput(features=f, additional_input_tokens=torch.zeros(2, 128, 1)))
    with pytest.raises(RuntimeError):  # CUDA only, no CPU fallback
        g(U.MultiViewTransformerInput(features=f))
    r = U.MultiViewGlobalAttentionTransformer(custom_positional_encoding="rope", **kw)
    assert isinstance(r.custom_positional_encoding, U.RoPE2D)
    with pytest.raises(ValueError):  # :341-351
        r(U.MultiViewTransformerInput(features=f, additional_input_tokens=torch.zeros(1, 128, 1)))
    with pytest.raises(ValueError):  # unknown positional-encoding name (:159-163)
        U.MultiViewGlobalAttentionTransformer(custom_positional_encoding="sincos", **kw)
    # option containers: qk_norm -> per-head LayerNorm(64); init_values -> LayerScale gamma; flags that are not built refuse
    o = U.MultiViewAlternatingAttentionTransformer(qk_norm=True, init_values=0.1, gradient_checkpointing=True, **kw)
    blk = o.self_attention_blocks[0]
    assert blk.attn.q_norm.weight.shape == (64,) and float(blk.ls2.gamma.detach()[0]) == pytest.approx(0.1)
    assert o.gradient_checkpointing
    with pytest.raises(NotImplementedError):
        U.MultiViewGlobalAttentionTransformer(drop_path=0.1, **kw)
    with pytest.raises(NotImplementedError):  # head_dim must be 64
        U.MultiViewGlobalAttentionTransformer(name="g", input_embed_dim=128, depth=1, dim=128, num_heads=4)
    ifr = U.MultiViewGlobalAttentionTransformerIFR(indices=[0], intermediates_only=True, **kw)
    assert ifr.indices == [0] and ifr.intermediates_only and ifr.norm_intermediate


def test_adaptors_match_oracle_on_cpu():
    import dust3r_oracle as O

    x = torch.randn(2, 4, 8, 6)
    ad = U.PointMapWithConfidenceAdaptor(name="pointmap", pointmap_mode="exp", pointmap_vmin=-float("inf"),
                                         pointmap_vmax=float("inf"), confidence_type="exp", confidence_vmin=1,
                                         confidence_vmax=float("inf"))
    out = ad(U.AdaptorInput(adaptor_feature=x, output_shape_hw=(8, 6)))
    p, c = O.pointmap_conf_adaptor(x)
    assert torch.allclose(out.value, p) and torch.allclose(out.confidence, c) and ad.fusable()
    d = U.DepthAdaptor(name="d", mode="exp", vmin=-float("inf"), vmax=float("inf"))
    assert torch.allclose(d(U.AdaptorInput(adaptor_feature=x[:, :1], output_shape_hw=(8, 6))).value, O.depth_adaptor(x[:, :1]))


def test_diff_attention_family_state_dict_matches_reference():
    """the DiffAttention transformer's parameter names, shapes and ORDER equal the reference's (drop-in checkpoints)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_import

    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    ref_import.import_reference()
    from uniception.models.info_sharing.diff_cross_attention_transformer import (
        DifferentialMultiViewCrossAttentionTransformer as RefDiff, DifferentialMultiViewCrossAttentionTransformerIFR as RefDiffIFR)
    from uniception.models.libs.croco.pos_embed import RoPE2D as RefRoPE

    import uniception_b200 as U

    kw = dict(name="d", input_embed_dim=192, num_views=3, depth=2, dim=256, num_heads=4)
    ours = U.DifferentialMultiViewCrossAttentionTransformer(custom_positional_encoding=U.RoPE2D(freq=100.0), **kw)
    ref = RefDiff(custom_positional_encoding=RefRoPE(freq=100.0), **kw)
    assert [(k, tuple(v.shape)) for k, v in ours.state_dict().items()] == [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    ours_ifr = U.DifferentialMultiViewCrossAttentionTransformerIFR(indices=[0, 1], **kw)
    ref_ifr = RefDiffIFR(indices=[0, 1], **kw)
    assert list(ours_ifr.state_dict()) == list(ref_ifr.state_dict())
    # lambda initialisation constants of the reference (transformer_blocks.py:682-683) per depth
    for i in range(2):
        assert abs(ours.multi_view_branches[0][i].cross_attn.lambda_init - ref.multi_view_branches[0][i].cross_attn.lambda_init) < 1e-12
    # head geometry: blocks are built with num_heads // 2 (diff_cross_attention_transformer.py:110-113)
    blk = ours.multi_view_branches[1][0]
    assert blk.attn.num_heads == 2 and blk.attn.head_dim == 128 and blk.cross_attn.num_heads == 2 and blk.cross_attn.head_dim == 64
    with pytest.raises(AssertionError):
        U.DifferentialMultiViewCrossAttentionTransformer(name="d", input_embed_dim=192, num_views=2, depth=1, dim=256, num_heads=3)
