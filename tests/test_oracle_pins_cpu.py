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


def test_oracle_scheduler_pinned_to_published_constants_and_torch_distributions():
    """Scheduler oracle (scheduling_ddim_flax.py:144-361) against things that are NOT this repo's code:
    * the published Stable Diffusion noise schedule (scaled_linear, beta 0.00085 .. 0.012, 1000 steps): alpha_bar_0 =
      1 - 0.00085, alpha_bar_999 = 0.0046601 (the `final` signal level quoted for SD), monotone;
    * the DDIM paper's eq. 12 / 16 restated here in float64 from the paper (x0-prediction + direction + noise, sigma_t =
      eta sqrt((1 - a_prev) / (1 - a_t)) sqrt(1 - a_t / a_prev)): mean and sigma of a step;
    * the step's log-prob == torch.distributions.Normal(mean, sigma).log_prob(x_prev) averaged over the sample's elements;
    * eta = 0 is deterministic and invertible in closed form (x_prev reproduces the paper's eq. 13 with sigma = 0)."""
    from oracle import scheduler as OS
    st = OS.set_timesteps(OS.SD_CONFIG, OS.create_state(OS.SD_CONFIG), 50)
    ac = np.asarray(st.alphas_cumprod, np.float64)
    assert abs(ac[0] - (1 - 0.00085)) < 1e-7 and abs(ac[999] - 0.0046601) < 2e-7 and np.all(np.diff(ac) < 0)
    assert list(st.timesteps[:3]) == [980 + OS.SD_CONFIG.steps_offset, 960 + OS.SD_CONFIG.steps_offset, 940 + OS.SD_CONFIG.steps_offset]
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    eps = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    xp = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    t = np.asarray([st.timesteps[0], st.timesteps[20], st.timesteps[49]])
    for eta in (1.0, 0.3):
        _, _, lp, mean = OS.step(OS.SD_CONFIG, st, eps, t, x, prev_sample=xp, eta=eta, return_mean=True)
        a_t = ac[t]
        prev_t = t - 1000 // 50
        a_p = np.where(prev_t >= 0, ac[np.maximum(prev_t, 0)], float(st.final_alpha_cumprod))
        sig = eta * np.sqrt((1 - a_p) / (1 - a_t)) * np.sqrt(1 - a_t / a_p)
        bc = lambda v: v.reshape(-1, 1, 1, 1)
        x0 = (x.astype(np.float64) - np.sqrt(1 - bc(a_t)) * eps) / np.sqrt(bc(a_t))
        ref_mean = np.sqrt(bc(a_p)) * x0 + np.sqrt(1 - bc(a_p) - bc(sig) ** 2) * eps
        assert np.abs(mean - ref_mean).max() < 5e-5 * max(1.0, np.abs(ref_mean).max())
        ref_lp = torch.distributions.Normal(torch.from_numpy(ref_mean), torch.from_numpy(bc(np.maximum(sig, 1e-6)))).log_prob(
            torch.from_numpy(xp.astype(np.float64))).mean(dim=(1, 2, 3)).numpy()
        assert np.abs(lp - ref_lp).max() < 2e-3 * np.abs(ref_lp).max() + 1e-4, (lp, ref_lp)
    prev, _, _ = OS.step(OS.SD_CONFIG, st, eps, t, x, prev_sample=None, key=(1, 2), eta=0.0)
    prev2, _, _ = OS.step(OS.SD_CONFIG, st, eps, t, x, prev_sample=None, key=(7, 9), eta=0.0)
    assert np.array_equal(prev, prev2)                                   # no noise enters at eta = 0


def test_oracle_ppo_loss_equals_torch_autograd_of_the_published_objective():
    """PPO clipped surrogate (reference training/policy_gradient.py:121-134; Schulman et al. 2017 eq. 7 with the sign
    flipped): the oracle's loss AND its analytic gradient against torch autograd of -min(r A, clip(r, 1-e, 1+e) A)."""
    from oracle import ppo
    rng = np.random.default_rng(2)
    lp = rng.normal(0, 0.2, 64).astype(np.float32)
    old = rng.normal(0, 0.2, 64).astype(np.float32)
    adv = (rng.normal(0, 4, 64)).astype(np.float32)          # some beyond the +-10 advantage clip after scaling
    adv[:4] = [15.0, -12.0, 0.0, 3.0]
    for clip in (1e-4, 0.2):
        loss, info, dlp = ppo.ppo_loss(lp, old, adv, clip)
        tlp = torch.tensor(lp, dtype=torch.float64, requires_grad=True)
        a = torch.tensor(adv, dtype=torch.float64).clamp(-10.0, 10.0)
        r = torch.exp(tlp - torch.tensor(old, dtype=torch.float64))
        obj = -torch.minimum(r * a, torch.clamp(r, 1 - clip, 1 + clip) * a).mean()
        obj.backward()
        assert abs(float(loss) - float(obj.detach())) < 1e-5 * max(1.0, abs(float(obj.detach())))
        assert np.abs(dlp - tlp.grad.numpy()).max() < 1e-5 * max(1.0, np.abs(tlp.grad.numpy()).max())
        assert abs(float(info["approx_kl"]) - 0.5 * float(((tlp.detach() - torch.tensor(old, dtype=torch.float64)) ** 2).mean())) < 1e-6
        assert abs(float(info["clipfrac"]) - float(((r.detach() - 1).abs() > clip).double().mean())) < 1e-6


def test_oracle_optimizer_equals_torch_adamw_with_global_norm_clipping():
    """`optax.chain(clip_by_global_norm, adamw)` restated in oracle/optim.py against torch.optim.AdamW +
    torch.nn.utils.clip_grad_norm_ (decoupled weight decay, bias-corrected moments, eps outside the square root: the two
    libraries implement the same update).  The reference keeps the first moment in bf16 (`mu_dtype`), torch in fp32: the
    comparison therefore bounds the difference by the bf16 rounding of mu (2^-8 relative on the update), and is exact to
    fp32 rounding on the first step, whose mu enters the update before it is rounded."""
    from oracle import optim as OO
    rng = np.random.default_rng(4)
    n = 4096
    p0 = rng.normal(0, 0.05, n).astype(np.float32)
    hp = dict(lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-2, max_norm=1.0)
    st = OO.AdamWState(n)
    tp = torch.nn.Parameter(torch.tensor(p0))
    opt = torch.optim.AdamW([tp], lr=hp["lr"], betas=(hp["b1"], hp["b2"]), eps=hp["eps"], weight_decay=hp["wd"])
    p = p0.copy()
    for it in range(6):
        g = (rng.normal(0, 1.0 if it % 2 == 0 else 1e-3, n)).astype(np.float32)   # alternately clipped / not clipped
        p, gn = OO.clip_adamw_update(p, g, st, **hp)
        tp.grad = torch.tensor(g)
        tn = torch.nn.utils.clip_grad_norm_([tp], hp["max_norm"])
        opt.step()
        assert abs(float(gn) - float(tn)) < 1e-4 * float(tn)
        step_size = np.abs(p - p0).max() if it == 0 else None
        err = np.abs(p - tp.detach().numpy()).max()
        if it == 0:
            assert err < 1e-6 * max(1.0, step_size / hp["lr"]), err              # fp32 rounding only
        assert err < (it + 1) * hp["lr"] * 2.0 ** -7, (it, err)                  # bf16 mu: < 2^-8 of an update per step
