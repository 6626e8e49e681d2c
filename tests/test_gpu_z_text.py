"""CLIP text encoder on the CUDA kernels (ddpo_b200/text_encoder.py, csrc/text.cu, causal flag of attention.cu) vs the
oracle restatement (oracle/text_encoder.py, itself pinned against transformers' CLIPTextModel on the CPU)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("b,heads,n", [(2, 2, 77), (1, 3, 128), (2, 1, 200), (1, 2, 5)])
def test_causal_attention(b, heads, n):
    from ddpo_b200 import ops
    g = torch.Generator().manual_seed(0)
    c = heads * 64
    qkv = bf(torch.randn(b * n, 3 * c, generator=g)).to(DEV)
    out = torch.zeros(b * n, c, dtype=torch.bfloat16, device=DEV)
    ops.attention_fwd(qkv, qkv[:, c:], qkv[:, 2 * c:], out, b, heads, n, n, 3 * c, 3 * c, 3 * c, c, causal=True)
    torch.cuda.synchronize()
    f = qkv.float().cpu().reshape(b, n, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = f[0] @ f[1].transpose(-1, -2) / 8.0 + torch.full((n, n), float("-inf")).triu(1)
    ref = (torch.softmax(s, -1) @ f[2]).permute(0, 2, 1, 3).reshape(b * n, c)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err < 2e-2, err
    # and the non-causal call is unchanged by the new flag
    out2 = torch.zeros_like(out)
    ops.attention_fwd(qkv, qkv[:, c:], qkv[:, 2 * c:], out2, b, heads, n, n, 3 * c, 3 * c, 3 * c, c)
    ref2 = (torch.softmax(f[0] @ f[1].transpose(-1, -2) / 8.0, -1) @ f[2]).permute(0, 2, 1, 3).reshape(b * n, c)
    torch.cuda.synchronize()
    assert (out2.float().cpu() - ref2).abs().max().item() < 2e-2


def test_text_kernels():
    from ddpo_b200 import ops
    g = torch.Generator().manual_seed(1)
    tok, pos = torch.randn(50, 64, generator=g).to(DEV), torch.randn(7, 64, generator=g).to(DEV)
    ids = torch.randint(0, 50, (21,), generator=g).to(DEV, torch.int32)
    out = torch.empty(21, 64, device=DEV)
    ops.embed_tokens(ids, tok, pos, out, 7)
    torch.cuda.synchronize()
    assert torch.equal(out, tok[ids.long()] + pos[torch.arange(21, device=DEV) % 7])
    x = (torch.randn(33, 256, generator=g) * 3).to(DEV)
    for act, fn in (("gelu", torch.nn.functional.gelu), ("quick_gelu", lambda t: t * torch.sigmoid(1.702 * t))):
        y = torch.empty(33, 256, dtype=torch.bfloat16, device=DEV)
        ops.act_bf16(x, y, act)
        torch.cuda.synchronize()
        ref = fn(x)
        assert ((y.float() - ref).abs() <= 5e-3 * ref.abs() + 1e-4).all()      # bf16 output: half an ulp = 0.4 %
    sc, bi = torch.randn(256, generator=g).to(DEV), torch.randn(256, generator=g).to(DEV)
    y = torch.empty(33, 256, device=DEV)
    ops.layernorm_f32(x, sc, bi, y, 33, 256, eps=1e-5)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (256,), sc, bi, 1e-5)
    np.testing.assert_allclose(y.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("cfg_name,b", [("TEXT_TINY", 2), ("TEXT_TINY_QUICK", 3)])
def test_text_encoder_matches_oracle(cfg_name, b):
    from ddpo_b200 import text_encoder as T
    from oracle import text_encoder as OT
    cfg = getattr(T, cfg_name)
    flat = T.init_flat_params(cfg, 0)
    enc = T.CLIPTextEncoder(cfg, flat, DEV)
    ids = torch.randint(3, cfg.vocab_size, (b, 77), generator=torch.Generator().manual_seed(5))
    got = enc(ids.numpy(), params=None, train=False)[0]
    torch.cuda.synchronize()
    ref = OT.encode(T.views(flat, cfg), cfg, ids)
    rel = ((got.cpu() - ref).norm() / ref.norm()).item()
    assert got.shape == (b, 77, cfg.hidden_size) and rel < 2e-2, rel
    # batch invariance + causality on the device path
    alone = enc(ids[1:2].numpy())[0]
    ids2 = ids.clone()
    ids2[:, 50] = (ids2[:, 50] + 1) % cfg.vocab_size
    later = enc(ids2.numpy())[0]
    torch.cuda.synchronize()
    assert torch.equal(alone[0], got[1])
    assert torch.equal(later[:, :50], got[:, :50]) and not torch.equal(later[:, 50:], got[:, 50:])


def test_text_encoder_full_size_runs():
    from ddpo_b200 import text_encoder as T
    enc = T.CLIPTextEncoder(T.SD2_TEXT, device=DEV, seed=0)
    ids = torch.randint(3, 49408, (8, 77), generator=torch.Generator().manual_seed(1))
    out = enc(ids.numpy())[0]
    torch.cuda.synchronize()
    assert out.shape == (8, 77, 1024) and torch.isfinite(out).all() and 0.5 < float(out.std()) < 2.0
