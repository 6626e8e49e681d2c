"""Pins of the U-Net oracle that need no JAX: ``oracle/unet.py`` (Flax semantics, NHWC, hand-written formulas) against
``tests/_torch_twin.py`` (the same network the PyTorch way, from torch.nn library primitives) on the same weights through
the standard Flax <-> PyTorch checkpoint conversion -- forward and parameter gradients, for the SD2-style (linear
projections, d_head 64) and the SD1-style (1x1-conv projections, 8 heads) topologies."""
import numpy as np
import pytest
import torch

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ddpo_b200 import unet_spec  # noqa: E402
from oracle.unet import UNetOracle
from _torch_twin import UNet2DConditionTwin, load_flax_params

SD1_TINY = unet_spec.UNetConfig(block_out_channels=(64, 128, 128, 128), attention_head_dim=(8, 8, 8, 8),
                                cross_attention_dim=96, use_linear_projection=False, sample_size=16, ctx_len=77)


@pytest.mark.parametrize("cfg", [unet_spec.TINY, SD1_TINY], ids=["sd2_style", "sd1_style"])
def test_oracle_unet_equals_pytorch_idiom_twin(cfg):
    flat = unet_spec.init_flat_params(cfg, 3)
    views = unet_spec.views(flat, cfg)
    twin = load_flax_params(UNet2DConditionTwin(cfg), views).double()
    g = torch.Generator().manual_seed(4)
    b, s = 2, cfg.sample_size
    x = torch.randn(b, 4, s, s, generator=g)
    ctx = torch.randn(b, cfg.ctx_len, cfg.cross_attention_dim, generator=g)
    ts = torch.tensor([981, 41])
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in views.items()}
    out_o = UNetOracle(cfg, p64, dtype=torch.float64)(x.double(), ts, ctx.double())
    out_t = twin(x.double(), ts, ctx.double())
    assert out_o.shape == out_t.shape == (b, 4, s, s)
    rel = ((out_o - out_t).norm() / out_t.norm()).item()
    assert rel < 1e-9, rel
    # fp32 oracle (what the tests and bench use) against the float64 twin: the size of fp32 round-off only
    out_o32 = UNetOracle(cfg, views)(x, ts, ctx)
    assert ((out_o32.double() - out_t).norm() / out_t.norm()).item() < 2e-5
    # parameter gradients of a scalar of the output
    w = torch.randn(out_t.shape, generator=g, dtype=torch.float64)
    (out_o * w).sum().backward()
    (out_t * w).sum().backward()
    tsd = dict(twin.named_parameters())
    from _torch_twin import flax_name_to_torch
    worst = 0.0
    for name, pv in p64.items():
        base, leaf = flax_name_to_torch(name)
        key = base + (".weight" if leaf in ("kernel", "scale") else ".bias")
        gt = tsd[key].grad
        if leaf == "kernel":
            gt = gt.permute(2, 3, 1, 0) if (gt.dim() == 4 and pv.dim() == 4) else gt.reshape(gt.shape[0], gt.shape[1]).t()
        go = pv.grad.reshape(gt.shape)
        worst = max(worst, ((go - gt).norm() / (gt.norm() + 1e-30)).item())
    assert worst < 1e-8, worst


def test_oracle_primitives_equal_torch_library_ops():
    """GroupNorm (Flax var = E[x^2] - E[x]^2 form), LayerNorm, tanh-GELU, SiLU and the attention core against torch's own."""
    import torch.nn.functional as F
    from oracle import unet as OU
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 8, 64, generator=g, dtype=torch.float64) * 3 + 1
    sc, bi = torch.randn(64, generator=g, dtype=torch.float64), torch.randn(64, generator=g, dtype=torch.float64)
    ref = F.group_norm(x.permute(0, 3, 1, 2), 32, sc, bi, 1e-5).permute(0, 2, 3, 1)
    assert (OU.group_norm(x, sc, bi) - ref).abs().max().item() < 1e-12
    assert (OU.layer_norm(x, sc, bi) - F.layer_norm(x, (64,), sc, bi, 1e-5)).abs().max().item() < 1e-12
    assert (OU.gelu_tanh(x) - F.gelu(x, approximate="tanh")).abs().max().item() < 1e-12
    assert (OU.silu(x) - F.silu(x)).abs().max().item() < 1e-12
    e = OU.timestep_embedding(torch.tensor([0, 1, 500, 999]), 320, torch.float64)
    k = torch.arange(160, dtype=torch.float64)
    fr = torch.tensor([0.0, 1.0, 500.0, 999.0], dtype=torch.float64)[:, None] * torch.exp(-np.log(10000.0) * k / 160)[None]
    # the oracle (like Flax) forms t * freq in fp32: arguments up to ~1e3 carry ~6e-5 of absolute error
    assert (e[:, :160] - torch.cos(fr)).abs().max().item() < 2e-4 and (e[:, 160:] - torch.sin(fr)).abs().max().item() < 2e-4


def test_oracle_vae_decoder_equals_pytorch_idiom_twin():
    """oracle/vae.py (Flax decoder, NHWC, hand-written single-head attention with the (C)^-1/4 double scaling) against the
    decoder written with torch library primitives on the same Flax-layout weights."""
    from _torch_twin import vae_decode_twin
    from ddpo_b200 import vae as V
    from oracle import vae as OV
    for cfg, seed in ((V.VAE_MICRO, 3), (V.VAE_MICRO, 4)):
        flat = V.init_flat_params(cfg, seed).double()
        views = V.views(flat, cfg)
        z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * 0.18215
        img_o, raw_o = OV.decode(views, cfg, z, dtype=torch.float64)
        img_t, raw_t = vae_decode_twin(views, cfg, z)
        assert raw_o.shape == raw_t.shape and img_o.shape == img_t.shape
        assert ((raw_o - raw_t).norm() / raw_t.norm()).item() < 1e-9
        assert (img_o - img_t).abs().max().item() < 1e-9


def test_oracle_vae_encoder_equals_pytorch_idiom_twin_and_has_the_published_size():
    from _torch_twin import vae_encode_twin
    from ddpo_b200 import vae as V
    from oracle import vae as OV
    # SD VAE: 83 653 863 parameters = decoder 49 490 199 + post_quant_conv 20 + encoder 34 163 592 + quant_conv 72
    assert V.num_params(V.SD_VAE, part="encoder") == 34_163_592 + 72
    cfg = V.VAE_MICRO
    flat = V.init_flat_params(cfg, 5, part="encoder").double()
    views = V.views(flat, cfg, part="encoder")
    img = torch.rand(2, 64, 64, 3, generator=torch.Generator().manual_seed(6), dtype=torch.float64)
    m_o = OV.encode(views, cfg, img, dtype=torch.float64)
    m_t = vae_encode_twin(views, cfg, img)
    assert m_o.shape == m_t.shape == (2, 8, 8, 8)
    assert ((m_o - m_t).norm() / m_t.norm()).item() < 1e-9
    # logvar clipping: scale the last layer up so that both bounds are hit
    views["quant_conv/kernel"].mul_(1e4)
    m = OV.encode(views, cfg, img, dtype=torch.float64)
    assert float(m[..., 4:].max()) == 20.0 and float(m[..., 4:].min()) == -30.0
