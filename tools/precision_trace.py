"""GPU: per-layer error trace of the encoder residual stream.

For ViT-B/16 @224 (BASELINE configs[0], the encoder whose output sat at 1.44x the autocast yardstick) and ViT-L/16 @512,
compares after EVERY block (un-normalised residual stream, through the intermediate-feature returner):
    ours (bf16 engine)  vs  fp32 oracle                      <- err_ours(layer)
    reference arithmetic under torch.autocast(bf16)  vs  fp32  <- err_autocast(layer)   (O.reference_functionals())
and the same for the final normalised output.  Prints one line per layer.

    python tools/precision_trace.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch

import dust3r_oracle as O
import uniception_b200 as U

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda"


def trace(C, depth, heads, S, seed=42):
    torch.manual_seed(seed)
    idx = list(range(depth))
    enc = U.CroCoIntermediateFeatureReturner(name="e", data_norm_type="dust3r", img_size=(S, S), enc_embed_dim=C, enc_depth=depth,
                                             enc_num_heads=heads, indices=idx, norm_intermediate=False, intermediates_only=False).to(DEV)
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(1, 3, S, S, generator=g).clamp_(-1, 1).to(DEV)
    with torch.no_grad():
        fin, inter = enc(U.ViTEncoderInput(image=img, data_norm_type="dust3r"))
        sd = {"encoder." + k: v.detach() for k, v in enc.state_dict().items()}
        with O.reference_functionals():
            f32, i32 = O.croco_encoder(sd, "encoder.", img, depth, heads, 16, indices=idx, norm_intermediate=False)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f16, i16 = O.croco_encoder(sd, "encoder.", img, depth, heads, 16, indices=idx, norm_intermediate=False)
    print(f"--- ViT C={C} depth={depth} heads={heads} {S}x{S}: rel-L2 error of the residual stream after each block")
    for i in range(depth):
        eo = O.parity(inter[i].features, i32[i])[1]
        el = O.parity(i16[i].float(), i32[i])[1]
        print(f"  block {i:2d}: ours {eo:.3e}  autocast-reference {el:.3e}  ratio {eo / max(el, 1e-12):.2f}")
    eo, el = O.parity(fin.features, f32)[1], O.parity(f16.float(), f32)[1]
    print(f"  final norm: ours {eo:.3e}  autocast-reference {el:.3e}  ratio {eo / max(el, 1e-12):.2f}")


if __name__ == "__main__":
    trace(768, 12, 12, 224)
    trace(1024, 24, 16, 512)
