"""Mirror of the reference's in-module self-test (info_sharing/cross_attention_transformer.py:515-609) on the B200 modules:
2 / 3 / 4 views with and without RoPE, IFR last-n / explicit indices / normalisation semantics -- plus a numeric check of the
3- and 4-view cross-attention against the oracle (Nk = (V-1)*N keys)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dust3r_oracle as O
import uniception_b200 as U

DEV = "cuda"
torch.manual_seed(0)
kw = dict(name="MV-CAT", input_embed_dim=192, depth=2, dim=128, num_heads=2)
ok = True
for rope in (None, U.RoPE2D(freq=100.0)):
    for nv in (2, 3, 4):
        m = U.MultiViewCrossAttentionTransformer(num_views=nv, custom_positional_encoding=rope, **kw).to(DEV)
        feats = [torch.rand(1, 192, 14, 14, device=DEV) for _ in range(nv)]
        out = m(U.MultiViewTransformerInput(features=feats))
        assert len(out.features) == nv and all(f.shape == (1, m.dim, 14, 14) for f in out.features)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        ref = O.info_sharing(sd, "", feats, 2, 2, base=100.0) if rope is not None else None
        if ref is not None:
            err = max(O.parity(out.features[v], ref[v])[1] for v in range(nv))
            print(f"views={nv} rope: rel {err:.3e}", flush=True)
            ok &= err <= 2e-2
m = U.MultiViewCrossAttentionTransformerIFR(num_views=2, indices=3, **{**kw, "depth": 4}).to(DEV)
x = U.MultiViewTransformerInput(features=[torch.rand(1, 192, 14, 14, device=DEV) for _ in range(2)])
o = m(x)
assert isinstance(o, tuple) and isinstance(o[0], U.MultiViewTransformerOutput) and len(o[1]) == 3 and len(o[1][0].features) == 2
m = U.MultiViewCrossAttentionTransformerIFR(num_views=2, indices=[0, 2], **{**kw, "depth": 4}).to(DEV)
o = m(x)
assert len(o[1]) == 2 and all(isinstance(t, U.MultiViewTransformerOutput) for t in o[1])
m = U.MultiViewCrossAttentionTransformerIFR(num_views=2, indices=[-1], norm_intermediate=False, **kw).to(DEV)
o = m(x)
assert all(not torch.equal(o[0].features[v], o[1][-1].features[v]) for v in range(2))
m = U.MultiViewCrossAttentionTransformerIFR(num_views=2, indices=[-1], norm_intermediate=True, **kw).to(DEV)
o = m(x)
assert all(torch.equal(o[0].features[v], o[1][-1].features[v]) for v in range(2))
print("SELFTEST", "OK" if ok else "BAD")
