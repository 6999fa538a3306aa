"""CPU: checkpoint compatibility.  A synthetic checkpoint with the ORIGINAL DUSt3R key names (the names
the reference's convert script reads, examples/models/dust3r/convert_dust3r_weights_to_uniception.py) is converted and
must load STRICTLY into our modules, with every tensor landing in the parameter the reference's mapping sends it to."""
import torch

import uniception_b200 as U
from uniception_b200 import checkpoints as CK
from uniception_b200.prediction_heads import DPTFeature, DPTRegressionProcessor

C_ENC, C_DEC, DEPTH_E, DEPTH_D, HE, HD = 128, 64, 2, 2, 2, 1


def _block(prefix, C, cross):
    g = torch.Generator().manual_seed(abs(hash(prefix)) % (2 ** 31))
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    sd = {f"{prefix}norm1.weight": r(C), f"{prefix}norm1.bias": r(C), f"{prefix}attn.qkv.weight": r(3 * C, C),
          f"{prefix}attn.qkv.bias": r(3 * C), f"{prefix}attn.proj.weight": r(C, C), f"{prefix}attn.proj.bias": r(C),
          f"{prefix}norm2.weight": r(C), f"{prefix}norm2.bias": r(C), f"{prefix}mlp.fc1.weight": r(4 * C, C),
          f"{prefix}mlp.fc1.bias": r(4 * C), f"{prefix}mlp.fc2.weight": r(C, 4 * C), f"{prefix}mlp.fc2.bias": r(C)}
    if cross:
        sd.update({f"{prefix}norm3.weight": r(C), f"{prefix}norm3.bias": r(C), f"{prefix}norm_y.weight": r(C), f"{prefix}norm_y.bias": r(C)})
        for n in ("projq", "projk", "projv", "proj"):
            sd[f"{prefix}cross_attn.{n}.weight"] = r(C, C)
            sd[f"{prefix}cross_attn.{n}.bias"] = r(C)
    return sd


def _original_checkpoint(two_decoders=True, head="linear"):
    g = torch.Generator().manual_seed(3)
    sd = {"patch_embed.proj.weight": torch.randn(C_ENC, 3, 16, 16, generator=g), "patch_embed.proj.bias": torch.randn(C_ENC, generator=g),
          "enc_norm.weight": torch.randn(C_ENC, generator=g), "enc_norm.bias": torch.randn(C_ENC, generator=g),
          "decoder_embed.weight": torch.randn(C_DEC, C_ENC, generator=g), "decoder_embed.bias": torch.randn(C_DEC, generator=g),
          "dec_norm.weight": torch.randn(C_DEC, generator=g), "dec_norm.bias": torch.randn(C_DEC, generator=g),
          "mask_token": torch.zeros(1, 1, C_DEC)}  # ignored by every converter
    for i in range(DEPTH_E):
        sd.update(_block(f"enc_blocks.{i}.", C_ENC, False))
    for i in range(DEPTH_D):
        sd.update(_block(f"dec_blocks.{i}.", C_DEC, True))
        if two_decoders:
            sd.update(_block(f"dec_blocks2.{i}.", C_DEC, True))
    for h in (1, 2):
        if head == "linear":
            sd[f"downstream_head{h}.proj.weight"] = torch.randn(4 * 256, C_DEC, generator=g)
            sd[f"downstream_head{h}.proj.bias"] = torch.randn(4 * 256, generator=g)
        else:
            feat = DPTFeature(patch_size=16, hooks=[0, 1, 2, 3], input_feature_dims=[C_ENC, C_DEC, C_DEC, C_DEC],
                              layer_dims=[12, 24, 48, 96], feature_dim=32)
            for k, v in feat.state_dict().items():
                sd[f"downstream_head{h}.dpt.{k}"] = torch.randn(v.shape, generator=g)
            for idx, shp in (("0", (16, 32, 3, 3)), ("2", (16, 16, 3, 3)), ("4", (4, 16, 1, 1))):
                sd[f"downstream_head{h}.dpt.head.{idx}.weight"] = torch.randn(*shp, generator=g)
                sd[f"downstream_head{h}.dpt.head.{idx}.bias"] = torch.randn(shp[0], generator=g)
    return {"model": sd}


def _model(head):
    return U.DUSt3R(name="t", img_size=(32, 32), pred_head_type=head, pred_head_feature_dim=32,
                    encoder_kwargs=dict(enc_embed_dim=C_ENC, enc_depth=DEPTH_E, enc_num_heads=HE),
                    info_sharing_kwargs=dict(depth=DEPTH_D, dim=C_DEC, num_heads=HD),
                    dpt_kwargs=dict(layer_dims=[12, 24, 48, 96]), dpt_indices=(0, 1))


def test_linear_checkpoint_loads_strictly_and_maps_tensors():
    ck = _original_checkpoint(True, "linear")
    m = _model("linear")
    res = m.load_state_dict(CK.dust3r_state_dict(ck, "linear"), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    o = ck["model"]
    assert torch.equal(m.encoder.enc_blocks[1].attn.qkv.weight, o["enc_blocks.1.attn.qkv.weight"])
    assert torch.equal(m.info_sharing.proj_embed.weight, o["decoder_embed.weight"])
    assert torch.equal(m.info_sharing.multi_view_branches[0][1].cross_attn.projk.weight, o["dec_blocks.1.cross_attn.projk.weight"])
    assert torch.equal(m.info_sharing.multi_view_branches[1][0].mlp.fc1.bias, o["dec_blocks2.0.mlp.fc1.bias"])
    assert torch.equal(m.info_sharing.norm.bias, o["dec_norm.bias"])
    assert torch.equal(m.head2.linear.weight[:, :, 0, 0], o["downstream_head2.proj.weight"])
    assert torch.equal(m.head1.linear.bias, o["downstream_head1.proj.bias"])


def test_croco_style_checkpoint_duplicates_the_decoder():
    ck = _original_checkpoint(False, "linear")  # only dec_blocks: both view branches get the same tensors (convert :28-34)
    sd = CK.cross_attention_state_dict(ck)
    m = _model("linear")
    assert set(sd) == set(m.info_sharing.state_dict())
    assert torch.equal(sd["multi_view_branches.0.1.attn.proj.weight"], sd["multi_view_branches.1.1.attn.proj.weight"])


def test_dpt_checkpoint_loads_strictly_and_maps_tensors():
    ck = _original_checkpoint(True, "dpt")
    m = _model("dpt")
    res = m.load_state_dict(CK.dust3r_state_dict(ck, "dpt"), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    o = ck["model"]
    assert torch.equal(m.dpt_regressor_head1.conv1.weight, o["downstream_head1.dpt.head.0.weight"])
    assert torch.equal(m.dpt_regressor_head2.conv2[2].bias, o["downstream_head2.dpt.head.4.bias"])
    assert torch.equal(m.dpt_feature_head2.scratch.refinenet3.out_conv.weight, o["downstream_head2.dpt.scratch.refinenet3.out_conv.weight"])
    f, p = CK.dpt_head_state_dicts(ck, 1)
    assert set(f) == set(DPTFeature(patch_size=16, hooks=[0, 1, 2, 3], input_feature_dims=[C_ENC, C_DEC, C_DEC, C_DEC],
                                    layer_dims=[12, 24, 48, 96], feature_dim=32).state_dict())
    assert set(p) == set(DPTRegressionProcessor(input_feature_dim=32, output_dim=4).state_dict())


def test_uniception_format_roundtrip(tmp_path):
    m = _model("linear")
    path = str(tmp_path / "ck.pth")
    torch.save({"model": m.state_dict(), "data_norm_type": "dust3r"}, path)
    m2 = _model("linear")
    res = CK.load_uniception_checkpoint(m2, path)
    assert not res.missing_keys and not res.unexpected_keys
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
