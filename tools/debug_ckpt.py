"""plain vs checkpointed alternating transformer: run-to-run and mode-to-mode differences of outputs / gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dust3r_oracle as O
import uniception_b200 as U

DEV = "cuda"
kw = dict(name="mv", input_embed_dim=192, depth=int(os.environ.get("DEPTH", 6)), dim=256, num_heads=4, use_rand_idx_pe_for_non_reference_views=False,
          custom_positional_encoding=U.RoPE2D(freq=100.0))
torch.manual_seed(3)
plain = U.MultiViewAlternatingAttentionTransformer(**kw).to(DEV)
ckpt = U.MultiViewAlternatingAttentionTransformer(gradient_checkpointing=True, **kw)
ckpt.load_state_dict(plain.state_dict())
ckpt = ckpt.to(DEV)
feats = [torch.randn(2, 192, 24, 24, device=DEV) for _ in range(3)]
cot = [torch.randn(2, 256, 24, 24, device=DEV) for _ in range(3)]
runs = []
for tag, m in (("plain", plain), ("plain", plain), ("ckpt", ckpt), ("ckpt", ckpt)):
    m.zero_grad(set_to_none=True)
    fin = [f.clone().requires_grad_(True) for f in feats]
    out = m(U.MultiViewTransformerInput(features=fin)).features
    sum((o * c).sum() for o, c in zip(out, cot)).backward()
    torch.cuda.synchronize()
    runs.append((tag, out[0].detach().clone(), fin[0].grad.clone(), fin[2].grad.clone(),
                 m.self_attention_blocks[2].attn.qkv.weight.grad.clone(), m.proj_embed.bias.grad.clone()))
names = ["out0", "d_in0", "d_in2", "d_qkv2", "d_proj_embed_b"]
for i in range(len(runs)):
    for j in range(i + 1, len(runs)):
        print(runs[i][0], i, "vs", runs[j][0], j, " ".join(f"{n}={O.parity(runs[i][k + 1], runs[j][k + 1])[1]:.2e}" for k, n in enumerate(names)))
