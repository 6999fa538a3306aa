"""uniception_b200 -- B200-native (sm_100a) implementation of UniCeption's DUSt3R two-view hot path.

Drop-in modules (same constructor signatures, dataclass I/O and state-dict keys as castacks/UniCeption):
    encoders.CroCoEncoder, encoders.CroCoIntermediateFeatureReturner
    info_sharing.MultiViewCrossAttentionTransformer(+IFR)
    prediction_heads.LinearFeature, PointMapWithConfidenceAdaptor, ...
    dust3r.DUSt3R
    rope.RoPE2D / cuRoPE2D  (the reference's one native op)
All compute goes through `libuc_b200.so` (C ABI in include/uc_b200.h); importing this package without
the built library raises ImportError -- there is no CPU or library fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is missing)
from . import checkpoints  # noqa: F401  (original DUSt3R / CroCo checkpoint -> UniCeption-format state dicts)
from .depth import ViTDPTDepth  # noqa: F401
from .dust3r import DUSt3R, interleave, is_symmetrized  # noqa: F401
from .encoders import (  # noqa: F401
    ENCODER_CONFIGS, CroCoEncoder, CroCoIntermediateFeatureReturner, IntermediateFeatureReturner, ManyAR_PatchEmbed, ViTEncoderInput, ViTEncoderOutput,
    encoder_factory, feature_returner_encoder_factory, feature_take_indices,
)
from .info_sharing import (  # noqa: F401
    INFO_SHARING_CLASSES, MultiViewAlternatingAttentionTransformer, MultiViewAlternatingAttentionTransformerIFR,
    MultiViewCrossAttentionTransformer, MultiViewCrossAttentionTransformerIFR, MultiViewGlobalAttentionTransformer,
    MultiViewGlobalAttentionTransformerIFR, MultiViewTransformerInput, MultiViewTransformerOutput,
)
from .prediction_heads import (  # noqa: F401
    AdaptorInput, ConfidenceAdaptor, DepthAdaptor, LinearFeature, PixelTaskOutput, PointMapAdaptor,
    PointMapWithConfidenceAdaptor, PredictionHeadInput,
)
from .rope import RoPE2D, cuRoPE2D  # noqa: F401

__version__ = "0.1.0"
from .diff_attention import (  # noqa: E402
    DiffAttention, DiffCrossAttention, DiffCrossAttentionBlock, DiffSelfAttentionBlock, DifferentialMultiViewCrossAttentionTransformer,
    DifferentialMultiViewCrossAttentionTransformerIFR, RMSNorm,
)
from . import registry  # noqa: E402,F401  (implementation switch / hook into the reference's registries; imports nothing from the reference)
