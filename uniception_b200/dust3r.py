"""Drop-in `DUSt3R` two-view model on the B200 engine (reference: uniception/models/factory/dust3r.py).

Same constructor, attributes, sub-module names (hence state-dict keys: `encoder.*`, `info_sharing.*`,
`head1.*`, `head2.*`), view-dict input and result dicts as the reference.  `forward` keeps the whole
pair in token-major bf16 between the three fused stages (no NCHW round trips) and runs the heads +
adaptor in fp32 like the reference's autocast-off region (dust3r.py:309).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn

from .checkpoints import load_checkpoint_file
from .encoders import CroCoEncoder, feature_take_indices
from .info_sharing import MultiViewCrossAttentionTransformer, MultiViewCrossAttentionTransformerIFR
from .params import ParamPack, get_pack
from .prediction_heads import (DPTFeature, DPTHead, DPTRegressionProcessor, LinearFeature, PointMapWithConfidenceAdaptor,
                               _HeadPostFn)
from .rope import RoPE2D


def is_symmetrized(gt1, gt2):
    "factory/dust3r.py:21-30: pairs (a, b) and (b, a) always both present"
    x = gt1["instance"]
    y = gt2["instance"]
    if len(x) == len(y) and len(x) == 1:
        return False  # special case of batchsize 1
    ok = True
    for i in range(0, len(x), 2):
        ok = ok and (x[i] == y[i + 1]) and (x[i + 1] == y[i])
    return ok


def interleave(tensor1, tensor2):
    "factory/dust3r.py:33-37"
    res1 = torch.stack((tensor1, tensor2), dim=1).flatten(0, 1)
    res2 = torch.stack((tensor2, tensor1), dim=1).flatten(0, 1)
    return res1, res2


class DUSt3R(nn.Module):
    "DUSt3R defined with UniCeption-compatible B200 modules"

    def __init__(
        self,
        name: str,
        data_norm_type: str = "dust3r",
        img_size: tuple = (224, 224),
        patch_embed_cls: str = "PatchEmbedDust3R",
        pred_head_type: str = "linear",
        pred_head_output_dim: int = 4,
        pred_head_feature_dim: int = 256,
        depth_mode: Tuple[str, float, float] = ("exp", -float("inf"), float("inf")),
        conf_mode: Tuple[str, float, float] = ("exp", 1, float("inf")),
        pos_embed: str = "RoPE100",
        pretrained_checkpoint_path: str = None,
        pretrained_encoder_checkpoint_path: str = None,
        pretrained_info_sharing_checkpoint_path: str = None,
        pretrained_pred_head_checkpoint_paths: List[str] = [None, None],
        pretrained_pred_head_regressor_checkpoint_paths: List[str] = [None, None],
        override_encoder_checkpoint_attributes: bool = False,
        # extension (not in the reference): override the hard-coded ViT-L / base-decoder sizes, used by tests
        encoder_kwargs: dict = None,
        info_sharing_kwargs: dict = None,
        dpt_kwargs: dict = None,
        dpt_indices: tuple = (5, 8),
        *args,
        **kwargs,
    ):
        super().__init__()
        self.name = name
        self.data_norm_type = data_norm_type
        self.img_size = img_size
        self.patch_embed_cls = patch_embed_cls
        self.pred_head_type = pred_head_type
        self.pred_head_output_dim = pred_head_output_dim
        self.depth_mode = depth_mode
        self.conf_mode = conf_mode
        self.pos_embed = pos_embed
        self.pretrained_checkpoint_path = pretrained_checkpoint_path

        freq = float(pos_embed[len("RoPE"):])
        self.rope = RoPE2D(freq=freq)

        self.encoder = CroCoEncoder(
            name=name, data_norm_type=data_norm_type, patch_embed_cls=patch_embed_cls, img_size=img_size,
            pretrained_checkpoint_path=pretrained_encoder_checkpoint_path,
            override_checkpoint_attributes=override_encoder_checkpoint_attributes, **(encoder_kwargs or {}))

        common = dict(name="base_info_sharing", input_embed_dim=self.encoder.enc_embed_dim, num_views=2,
                      custom_positional_encoding=self.rope,
                      pretrained_checkpoint_path=pretrained_info_sharing_checkpoint_path, **(info_sharing_kwargs or {}))
        if self.pred_head_type == "linear":
            self.info_sharing = MultiViewCrossAttentionTransformer(**common)
        elif self.pred_head_type == "dpt":
            # intermediate (un-normalised) decoder features of blocks 5 and 8 + the final ones (dust3r.py:135-144)
            self.info_sharing = MultiViewCrossAttentionTransformerIFR(indices=list(dpt_indices), norm_intermediate=False, **common)
        else:
            raise ValueError(f"Invalid prediction head type: {pred_head_type}. Must be 'linear' or 'dpt'.")

        if self.pred_head_type == "linear":
            self.head1 = LinearFeature(input_feature_dim=self.info_sharing.dim, output_dim=pred_head_output_dim,
                                       patch_size=self.encoder.patch_size,
                                       pretrained_checkpoint_path=pretrained_pred_head_checkpoint_paths[0])
            self.head2 = LinearFeature(input_feature_dim=self.info_sharing.dim, output_dim=pred_head_output_dim,
                                       patch_size=self.encoder.patch_size,
                                       pretrained_checkpoint_path=pretrained_pred_head_checkpoint_paths[1])
        else:
            dims = [self.encoder.enc_embed_dim] + [self.info_sharing.dim] * 3
            for k in (1, 2):  # dust3r.py:164-192: dpt_feature_head{k}, dpt_regressor_head{k}, head{k} = Sequential(both)
                f = DPTFeature(patch_size=self.encoder.patch_size, hooks=[0, 1, 2, 3], input_feature_dims=dims,
                               feature_dim=pred_head_feature_dim, pretrained_checkpoint_path=pretrained_pred_head_checkpoint_paths[k - 1],
                               **(dpt_kwargs or {}))
                r = DPTRegressionProcessor(input_feature_dim=pred_head_feature_dim, output_dim=pred_head_output_dim,
                                           pretrained_checkpoint_path=pretrained_pred_head_regressor_checkpoint_paths[k - 1])
                setattr(self, f"dpt_feature_head{k}", f)
                setattr(self, f"dpt_regressor_head{k}", r)
                setattr(self, f"head{k}", DPTHead(f, r))

        self.adaptor = PointMapWithConfidenceAdaptor(
            name="pointmap", pointmap_mode=depth_mode[0], pointmap_vmin=depth_mode[1], pointmap_vmax=depth_mode[2],
            confidence_type=conf_mode[0], confidence_vmin=conf_mode[1], confidence_vmax=conf_mode[2])
        if not self.adaptor.fusable() or pred_head_output_dim != 4:
            raise NotImplementedError("uniception_b200: only depth_mode=('exp',-inf,inf) / conf_mode=('exp',vmin,vmax) heads are fused")

        if self.pretrained_checkpoint_path is not None:
            print(f"Loading pretrained DUSt3R weights from {self.pretrained_checkpoint_path} ...")
            ckpt = load_checkpoint_file(self.pretrained_checkpoint_path)
            print(self.load_state_dict(ckpt["model"]))

    def pack(self) -> ParamPack:
        """Flat fp32 master / gradient / bf16 operand buffers of the whole model (params.py)."""
        return get_pack(self)

    def _encode_tokens(self, img1, img2, pk):
        "dust3r.py:211-225, same-shape branch: one encoder pass over the concatenated views"
        if img1.shape[-2:] != img2.shape[-2:]:
            raise NotImplementedError("uniception_b200: mixed-resolution view pairs (SURVEY.md 8f3; broken in the reference too, dust3r.py:221)")
        tok, _ = self.encoder.forward_tokens(torch.cat((img1, img2), dim=0), pk, "encoder.")
        return tok

    def forward(self, view1, view2):
        img1, img2 = view1["img"], view2["img"]
        _, _, height1, width1 = img1.shape
        self.encoder._check_data_normalization_type(view1["data_norm_type"])
        pk = self.pack()
        pk.refresh_bf16()  # one cast kernel: fp32 masters -> bf16 tensor-core operands

        p = self.encoder.patch_size
        h, w = height1 // p, width1 // p
        if is_symmetrized(view1, view2):
            # encode each unordered pair once, then interleave (dust3r.py:227-238)
            tok = self._encode_tokens(img1[::2], img2[::2], pk)
            Bh = img1[::2].shape[0]
            C = tok.shape[-1]
            f1, f2 = tok[: Bh * h * w].view(Bh, h * w, C), tok[Bh * h * w:].view(Bh, h * w, C)
            f1, f2 = interleave(f1, f2)
            t1, t2 = f1.reshape(-1, C), f2.reshape(-1, C)
        else:
            tok = self._encode_tokens(img1, img2, pk)
            half = tok.shape[0] // 2
            t1, t2 = tok[:half], tok[half:]
        B = img1.shape[0]
        cmin = float(self.adaptor.confidence_adaptor.vmin)
        cmax = float(self.adaptor.confidence_adaptor.vmax)
        if self.pred_head_type == "linear":
            (d1, d2), _ = self.info_sharing.forward_tokens([t1, t2], B, h, w, pk, "info_sharing.")
            pts1, conf1 = self.head1.forward_fused(d1, B, h, w, pk, "head1.", cmin, cmax)
            pts2, conf2 = self.head2.forward_fused(d2, B, h, w, pk, "head2.", cmin, cmax)
        else:
            take, _ = feature_take_indices(self.info_sharing.depth, self.info_sharing.indices)
            (d1, d2), inter = self.info_sharing.forward_tokens([t1, t2], B, h, w, pk, "info_sharing.", take, False)
            hw = (height1, width1)
            y1 = self.head1.forward_tokens([t1, inter[0][0], inter[1][0], d1], B, h, w, hw)  # dust3r.py:293-306
            y2 = self.head2.forward_tokens([t2, inter[0][1], inter[1][1], d2], B, h, w, hw)
            pts1, conf1 = _HeadPostFn.apply(y1, B, height1, width1, cmin, cmax)
            pts2, conf2 = _HeadPostFn.apply(y2, B, height1, width1, cmin, cmax)
        res1 = {"pts3d": pts1, "conf": conf1}
        res2 = {"pts3d_in_other_view": pts2, "conf": conf2}
        return res1, res2
