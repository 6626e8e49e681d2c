"""CPU tests of the oracle: known-answer pins, golden fixtures, self-consistency identities."""
import os

import numpy as np
import torch

from oracle import optim as OO, ppo, scheduler as S, threefry as T

G = os.path.join(os.path.dirname(__file__), "golden")


def test_threefry_random123_known_answers():
    kat = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, exp in kat:
        y0, y1 = T.threefry2x32_pair(key[0], key[1], np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
        assert (int(y0[0]), int(y1[0])) == exp


def test_jax_published_values():
    k = T.PRNGKey(0)
    assert k.tolist() == [0, 0]
    assert T.split(k).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert abs(float(T.normal(k, ())) - (-0.20584226)) < 1e-7
    assert abs(float(T.uniform(k, ())) - 0.41845703) < 1e-7
    assert abs(float(T.normal(T.split(k)[1], (1,))[0]) - (-1.2515389)) < 1e-6
    assert T.PRNGKey((5 << 32) + 7).tolist() == [5, 7]


def test_golden_prng_and_odd_sizes():
    g = np.load(os.path.join(G, "prng.npz"))
    keys = T.split(g["key"], 4)
    np.testing.assert_array_equal(keys, g["split4"])
    np.testing.assert_array_equal(T.normal(keys[1], (257,)), g["normal_257"])
    np.testing.assert_array_equal(T.normal(keys[2], (2, 4, 8, 8)), g["normal_2x4x8x8"])
    assert np.isfinite(g["normal_257"]).all() and abs(g["normal_2x4x8x8"].mean()) < 0.2


def test_scheduler_tables_and_golden_steps():
    st = S.set_timesteps(S.SD_CONFIG, S.create_state(S.SD_CONFIG), 50)
    assert st.timesteps[0] == 981 and st.timesteps[-1] == 1 and len(st.timesteps) == 50
    assert abs(st.alphas_cumprod[0] - 0.99915) < 1e-6 and abs(float(st.final_alpha_cumprod) - 0.99915) < 1e-6
    g = np.load(os.path.join(G, "ddim.npz"))
    eps = ppo.cfg_combine(g["eu"], g["ec"], 5.0)
    for t in (981, 501, 21, 1):
        prev, _, lp = S.step(S.SD_CONFIG, st, eps, t, g["x"], key=g["key"], eta=1.0)
        np.testing.assert_array_equal(prev, g[f"prev_{t}"])
        np.testing.assert_array_equal(lp, g[f"logp_{t}"])


def test_scheduler_identities():
    st = S.set_timesteps(S.SD_CONFIG, S.create_state(S.SD_CONFIG), 50)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4, 8, 8)).astype(np.float32)
    e = rng.standard_normal((2, 4, 8, 8)).astype(np.float32)
    # eta = 0: sigma clamps to 1e-6 for the density, prev == mean
    prev, _, lp, mean = S.step(S.SD_CONFIG, st, e, 501, x, key=T.PRNGKey(0), eta=0.0, return_mean=True)
    np.testing.assert_array_equal(prev, mean)
    np.testing.assert_allclose(lp, -np.log(1e-6) - 0.5 * np.log(2 * np.pi), rtol=1e-6)
    # sample then score the same x_prev: identical log-prob
    prev, _, lp = S.step(S.SD_CONFIG, st, e, 21, x, key=T.PRNGKey(1), eta=1.0)
    _, _, lp2 = S.step(S.SD_CONFIG, st, e, np.array([21, 21]), x, prev_sample=prev, eta=1.0)
    np.testing.assert_array_equal(lp, lp2)
    # both key and prev_sample / missing set_timesteps raise like the reference
    import pytest
    with pytest.raises(ValueError):
        S.step(S.SD_CONFIG, st, e, 21, x, key=T.PRNGKey(1), prev_sample=prev)
    with pytest.raises(ValueError):
        S.step(S.SD_CONFIG, S.create_state(S.SD_CONFIG), e, 21, x, key=T.PRNGKey(1))
    # finite-difference check of d log_prob / d eps
    dl = np.array([1.0, -2.0], np.float32)
    ga = S.logprob_grad_eps(S.SD_CONFIG, st, e, np.array([21, 21]), x, prev, 1.0, dl)
    idx = (1, 2, 3, 4)
    h = 1e-3
    ep, em = e.astype(np.float64).copy(), e.astype(np.float64).copy()
    ep[idx] += h
    em[idx] -= h

    def lp64(eps):
        a_t, a_prev, sig = (v.astype(np.float64) for v in S.coefficients(S.SD_CONFIG, st, np.array([21, 21]), 1.0))
        a_t, a_prev, sig = a_t[:, None, None, None], a_prev[:, None, None, None], sig[:, None, None, None]
        m = np.sqrt(a_prev) * (x - np.sqrt(1 - a_t) * eps) / np.sqrt(a_t) + np.sqrt(1 - a_prev - sig ** 2) * eps
        return (-(prev - m) ** 2 / (2 * sig ** 2) - np.log(sig) - 0.5 * np.log(2 * np.pi)).reshape(2, -1).mean(1)
    fd = ((lp64(ep) - lp64(em)) / (2 * h) * dl)[idx[0]]
    assert abs(fd - ga[idx]) < 1e-6 + 1e-4 * abs(fd)


def test_ppo_identities_and_golden():
    lp = np.array([0.3, -1.2, 0.5], np.float32)
    adv = np.array([2.0, -30.0, 0.5], np.float32)
    loss, info, dlp = ppo.ppo_loss(lp, lp, adv, 1e-4)
    assert info["approx_kl"] == 0 and info["clipfrac"] == 0
    np.testing.assert_allclose(loss, -np.mean(np.clip(adv, -10, 10)), rtol=1e-6)
    np.testing.assert_allclose(dlp, -np.clip(adv, -10, 10) / 3, rtol=1e-6)
    g = np.load(os.path.join(G, "ppo.npz"))
    loss, info, dlp = ppo.ppo_loss(g["lp"], g["old"], g["adv"], 1e-4)
    np.testing.assert_array_equal(loss, g["loss"])
    np.testing.assert_array_equal(dlp, g["dlp"])
    # autograd cross-check of the analytic gradient
    t = torch.tensor(g["lp"], requires_grad=True)
    a = torch.clamp(torch.tensor(g["adv"]), -10, 10)
    r = torch.exp(t - torch.tensor(g["old"]))
    torch.maximum(-a * r, -a * torch.clamp(r, 1 - 1e-4, 1 + 1e-4)).mean().backward()
    np.testing.assert_allclose(dlp, t.grad.numpy(), rtol=1e-5, atol=1e-8)


def test_adamw_one_step_closed_form():
    p = np.array([1.0, -2.0, 0.5, 3.0], np.float32)
    g = np.array([0.1, -0.2, 0.3, 0.05], np.float32)
    st = OO.AdamWState(4)
    new, gn = OO.clip_adamw_update(p, g, st, lr=1e-2, wd=1e-4)
    # first step: mu_hat = g, nu_hat = g^2 -> update = sign(g) (up to eps) + wd p ; no clipping (|g| < 1)
    np.testing.assert_allclose(new, p - 1e-2 * (np.sign(g) + 1e-4 * p), rtol=1e-5)
    assert abs(gn - np.linalg.norm(g)) < 1e-7
    # clipping: norm 10 -> scaled to 1; mu stored in bf16
    st = OO.AdamWState(4)
    new, gn = OO.clip_adamw_update(p, 50 * g, st, lr=1e-2, wd=0.0)
    np.testing.assert_allclose(st.mu, OO.to_bf16_f32(0.1 * 50 * g / gn), rtol=1e-6)
    # accumulate-then-update == mean of micro-gradients
    a = OO.AccumulatingTrainState(p, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, max_norm=1e9)
    a.apply_gradients(g, False)
    a.apply_gradients(3 * g, True)
    b = OO.AccumulatingTrainState(p, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, max_norm=1e9)
    b.apply_gradients(2 * g, True)
    np.testing.assert_allclose(a.params, b.params, rtol=1e-6)


def test_unet_oracle_golden_and_param_count():
    from ddpo_b200 import unet_spec
    from oracle.unet import UNetOracle
    assert unet_spec.num_params(unet_spec.SD2_BASE) == 865_910_724   # SD2-base U-Net
    assert unet_spec.num_params(unet_spec.SD1) == 859_520_964        # SD v1.x U-Net
    g = np.load(os.path.join(G, "unet_tiny.npz"))
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 0)
    assert abs(flat.double().sum().item() - float(g["param_sum"])) < 1e-6
    y = UNetOracle(cfg, unet_spec.views(flat, cfg))(torch.from_numpy(g["lat"]), torch.from_numpy(g["ts"]),
                                                    torch.from_numpy(g["ctx"]))
    np.testing.assert_allclose(y.numpy(), g["eps"], rtol=1e-4, atol=1e-5)
    y64 = UNetOracle(cfg, unet_spec.views(flat, cfg), torch.float64)(torch.from_numpy(g["lat"]),
                                                                      torch.from_numpy(g["ts"]), torch.from_numpy(g["ctx"]))
    assert (y64.float() - y).abs().max().item() < 1e-4


def test_scheduler_prediction_types_are_consistent():
    """`v_prediction` (scheduling_ddim_flax.py:309-316) is the epsilon branch fed eps = sqrt(a) v + sqrt(1-a) x; the
    analytic gradient of every branch agrees with finite differences; unknown types raise the reference's ValueError."""
    from dataclasses import replace
    import pytest
    from oracle import scheduler as S
    st = S.set_timesteps(S.SD_CONFIG, S.create_state(S.SD_CONFIG), 50)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 4, 8, 8)).astype(np.float32)
    v = rng.standard_normal((2, 4, 8, 8)).astype(np.float32)
    nxt = rng.standard_normal((2, 4, 8, 8)).astype(np.float32)
    ts = np.array([981, 301])
    a = st.alphas_cumprod[ts].reshape(2, 1, 1, 1)
    eps = np.sqrt(a) * v + np.sqrt(1 - a) * x
    vcfg = replace(S.SD_CONFIG, prediction_type="v_prediction")
    _, _, lp_v, mean_v = S.step(vcfg, st, v, ts, x, prev_sample=nxt, eta=1.0, return_mean=True)
    _, _, lp_e, mean_e = S.step(S.SD_CONFIG, st, eps, ts, x, prev_sample=nxt, eta=1.0, return_mean=True)
    np.testing.assert_allclose(mean_v, mean_e, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(lp_v, lp_e, rtol=1e-4)
    for pred in ("epsilon", "sample", "v_prediction"):
        cfg = replace(S.SD_CONFIG, prediction_type=pred)
        dl = np.array([1.0, -2.0])
        g = S.logprob_grad_eps(cfg, st, v, ts, x, nxt, 1.0, dl)
        d = rng.standard_normal(v.shape)
        h = 1e-3
        f = lambda m: float((S.step(cfg, st, m.astype(np.float64), ts, x, prev_sample=nxt, eta=1.0)[2].astype(np.float64) * dl).sum())
        # step() computes in fp32: central difference with a step large enough for fp32 round-off
        num = (f(v + h * d) - f(v - h * d)) / (2 * h)
        np.testing.assert_allclose((g * d).sum(), num, rtol=5e-2)
    with pytest.raises(ValueError):
        S.step(replace(S.SD_CONFIG, prediction_type="velocity"), st, v, ts, x, prev_sample=nxt, eta=1.0)
