"""CPU: implementation switch / registry hook (uniception_b200/registry.py).
The hook half needs the real reference importable (this container only; skipped on the GPU box)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import uniception_b200 as U
from uniception_b200 import registry

try:
    import ref_import

    HAVE_REF = ref_import.reference_available()
except Exception:  # pragma: no cover
    HAVE_REF = False


def test_b200_lookup_needs_no_reference():
    assert registry.encoder_class("croco") is U.CroCoEncoder
    assert registry.encoder_class("croco", feature_returner=True) is U.CroCoIntermediateFeatureReturner
    assert registry.info_sharing_class("cross_attention") is U.MultiViewCrossAttentionTransformer
    assert registry.info_sharing_class("global_attention", ifr=True) is U.MultiViewGlobalAttentionTransformerIFR
    with pytest.raises(ValueError):
        registry.encoder_class("dinov2")  # not part of the B200 path
    assert registry.info_sharing_class("diff_cross_attention") is U.DifferentialMultiViewCrossAttentionTransformer
    with pytest.raises(ValueError):
        registry.info_sharing_class("no_such_attention")
    with pytest.raises(ValueError):
        registry.encoder_class("croco", implementation="tpu")
    enc = registry.build_encoder("croco", name="e", data_norm_type="dust3r", enc_embed_dim=128, enc_depth=1, enc_num_heads=2)
    assert isinstance(enc, U.CroCoEncoder)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_install_swaps_the_reference_registries_and_restores_them():
    ref_import.import_reference()
    import uniception.models.encoders as RE
    import uniception.models.info_sharing as RI

    ref_croco, ref_cross = RE.ENCODER_CONFIGS["croco"]["class"], RI.INFO_SHARING_CLASSES["cross_attention"]
    ref_dino = RE.ENCODER_CONFIGS["dinov2"]["class"]
    assert registry.current() == "reference"
    with registry.implementation("b200"):
        assert registry.current() == "b200"
        # the reference's OWN factory now builds the B200 module, with the reference's keyword arguments
        enc = RE.encoder_factory("croco", name="e", data_norm_type="dust3r", enc_embed_dim=128, enc_depth=1, enc_num_heads=2)
        assert isinstance(enc, U.CroCoEncoder)
        ifr = RE.feature_returner_encoder_factory("croco", name="e", data_norm_type="dust3r", enc_embed_dim=128, enc_depth=2,
                                                  enc_num_heads=2, indices=[0, 1])
        assert isinstance(ifr, U.CroCoIntermediateFeatureReturner)
        assert RE.ENCODER_CONFIGS["croco"]["supported_models"] == ["CroCov2", "DUSt3R", "MASt3R"]  # metadata kept
        assert RI.INFO_SHARING_CLASSES["cross_attention"][0] is U.MultiViewCrossAttentionTransformer
        assert RI.INFO_SHARING_CLASSES["alternating_attention"][1] is U.MultiViewAlternatingAttentionTransformerIFR
        assert RI.INFO_SHARING_CLASSES["diff_cross_attention"][0] is U.DifferentialMultiViewCrossAttentionTransformer
        # untouched names keep the reference's classes
        assert RE.ENCODER_CONFIGS["dinov2"]["class"] is ref_dino
        # explicit per-module choice still reaches the reference while the hook is on
        assert registry.encoder_class("croco", implementation="reference") is ref_croco
        assert registry.info_sharing_class("cross_attention", implementation="reference") is ref_cross[0]
        # same state-dict keys and shapes either way (drop-in)
        ref_enc = ref_croco(name="e", data_norm_type="dust3r", enc_embed_dim=128, enc_depth=1, enc_num_heads=2)
        assert {k: tuple(v.shape) for k, v in ref_enc.state_dict().items()} == {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert registry.current() == "reference"
    assert RE.ENCODER_CONFIGS["croco"]["class"] is ref_croco and RI.INFO_SHARING_CLASSES["cross_attention"] is ref_cross
    registry.install("b200")
    registry.uninstall()
    assert RE.ENCODER_CONFIGS["croco"]["class"] is ref_croco
