"""Generates the second batch of fixtures (python tests/golden/make_golden_r1b.py): RWR inputs / loss, VAE decoder,
CLIP text encoder.

* ``text_tiny.npz`` is produced by the installed ``transformers`` library's own ``CLIPTextModel`` (the PyTorch twin of
  the ``FlaxCLIPTextModel`` the reference loads) -- a fixture from a LIVE third-party implementation, not from the oracle.
* ``rwr.npz`` and ``vae_micro.npz`` pin the oracle's own behaviour (no JAX / diffusers offline): regression pins that the
  CPU suite checks against the oracle and the GPU suite against the CUDA path."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ddpo_b200 import text_encoder as T, vae as V  # noqa: E402
from oracle import diffusion as OD, scheduler as OS, threefry, vae as OV  # noqa: E402


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    # ---- RWR: key lineage, posterior sample, noise, timesteps, add_noise, weighted CFG-MSE ----
    g = torch.Generator().manual_seed(0)
    mom = torch.randn(3, 8, 8, 8, generator=g).numpy()
    mom[..., 4:] = mom[..., 4:] * 2 - 3
    train_rng = threefry.PRNGKey(2024)
    _, sample_rng, new_rng = OD.split3(train_rng)
    ac = OS.create_state(OS.SD_CONFIG).alphas_cumprod
    noisy, noise, ts, lat = OD.make_inputs(mom, sample_rng, ac)
    eu = torch.randn(3, 256, generator=g)
    ec = torch.randn(3, 256, generator=g)
    w = np.array([0.5, 0.3, 0.2], np.float32)
    loss, per = OD.mse_loss(eu, ec, torch.from_numpy(noise.reshape(3, -1)), 5.0, True, w)
    np.savez(os.path.join(out, "rwr.npz"), moments=mom, train_rng=train_rng, sample_rng=sample_rng, new_rng=new_rng,
             noisy=noisy, noise=noise, timesteps=ts, latents=lat, eps_u=eu.numpy(), eps_c=ec.numpy(), weights=w,
             loss=np.float32(loss.item()), per_sample=per.numpy(),
             randint_17=threefry.randint(threefry.PRNGKey(5), (17,), 0, 1000))
    # ---- VAE decoder (VAE_MICRO, 8x8 latents -> 64x64 images) ----
    cfg = V.VAE_MICRO
    flat = V.init_flat_params(cfg, 0)
    z = (torch.randn(2, 4, 8, 8, generator=g) * 0.18215).numpy()
    img, raw = OV.decode(V.views(flat, cfg), cfg, z)
    np.savez(os.path.join(out, "vae_micro.npz"), latents=z, raw=raw.numpy(), image_mean=img.mean(dim=(1, 2)).numpy(),
             param_sum=np.float64(flat.double().sum().item()))
    # ---- CLIP text encoder: transformers' own implementation ----
    import test_text_encoder_cpu as TT
    for name, tcfg in (("gelu", T.TEXT_TINY), ("quick_gelu", T.TEXT_TINY_QUICK)):
        tflat = T.init_flat_params(tcfg, 0)
        model = TT._hf_model(tcfg, tflat)
        ids = torch.randint(3, tcfg.vocab_size, (2, 77), generator=g)
        with torch.no_grad():
            ref = model(input_ids=ids).last_hidden_state.numpy()
        np.savez(os.path.join(out, f"text_tiny_{name}.npz"), input_ids=ids.numpy(), last_hidden_state=ref,
                 param_sum=np.float64(tflat.double().sum().item()))


if __name__ == "__main__":
    main()
