"""Generates tests/golden/*.npz from the CPU oracle (run: python tests/golden/make_golden.py).

The reference ships no fixtures and JAX/Flax/diffusers cannot be imported here, so these vectors pin the
ORACLE's own behaviour (regression pins, checked on CPU and against the CUDA path on the GPU).  The PRNG
entries additionally coincide with published JAX outputs (see tests/test_oracle_prng.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scheduler as S, threefry as T  # noqa: E402
from oracle import ppo  # noqa: E402


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    key = T.PRNGKey(1234)
    keys = T.split(key, 4)
    np.savez(os.path.join(out, "prng.npz"), key=key, split4=keys, normal_257=T.normal(keys[1], (257,)),
             normal_2x4x8x8=T.normal(keys[2], (2, 4, 8, 8)), uniform_16=T.uniform(keys[3], (16,)))
    st = S.set_timesteps(S.SD_CONFIG, S.create_state(S.SD_CONFIG), 50)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    eu = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    ec = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    eps = ppo.cfg_combine(eu, ec, 5.0)
    res = {}
    for t in (981, 501, 21, 1):
        prev, _, lp = S.step(S.SD_CONFIG, st, eps, t, x, key=keys[0], eta=1.0)
        res[f"prev_{t}"], res[f"logp_{t}"] = prev, lp
    ts = np.array([981, 21, 1])
    _, _, lpv = S.step(S.SD_CONFIG, st, eps, ts, x, prev_sample=res["prev_501"], eta=1.0)
    np.savez(os.path.join(out, "ddim.npz"), x=x, eu=eu, ec=ec, key=keys[0], alphas_cumprod=st.alphas_cumprod,
             timesteps=st.timesteps, logp_vec_ts=lpv, **res)
    lp = rng.standard_normal(8).astype(np.float32)
    old = (lp + 2e-4 * rng.standard_normal(8)).astype(np.float32)
    adv = (12 * rng.standard_normal(8)).astype(np.float32)
    loss, info, dlp = ppo.ppo_loss(lp, old, adv, 1e-4)
    np.savez(os.path.join(out, "ppo.npz"), lp=lp, old=old, adv=adv, loss=loss, approx_kl=info["approx_kl"],
             clipfrac=info["clipfrac"], dlp=dlp)
    from ddpo_b200 import unet_spec
    from oracle.unet import UNetOracle
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    y = UNetOracle(cfg, unet_spec.views(flat, cfg))(lat, torch.tensor([981, 21]), ctx)
    np.savez(os.path.join(out, "unet_tiny.npz"), lat=lat.numpy(), ctx=ctx.numpy(), ts=np.array([981, 21]),
             eps=y.numpy(), param_sum=np.float64(flat.double().sum().item()))


if __name__ == "__main__":
    main()
