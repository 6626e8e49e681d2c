"""CPU restatement of the reward-weighted-regression train step ``ddpo/training/diffusion.py:6-102``
(posterior sample of the stored VAE moments -> forward diffusion -> U-Net (cond [+ uncond]) -> MSE on the
CFG-combined prediction, optionally weighted).  TEST INFRASTRUCTURE ONLY.

3P pieces restated: diffusers==0.12.1 ``vae_flax.FlaxDiagonalGaussianDistribution`` (mean/logvar split on the last
axis, logvar clipped to [-30, 20], ``sample = mean + exp(0.5 logvar) * normal(key, mean.shape)``) and
``scheduling_utils_flax.add_noise_common`` (``sqrt(a_t) x + sqrt(1 - a_t) n``)."""
import numpy as np
import torch

from . import threefry

VAE_SCALING = np.float32(0.18215)


def split3(train_rng):
    """``dropout_rng, sample_rng, new_train_rng = jax.random.split(train_rng, 3)`` (:14)."""
    k = threefry.split(np.asarray(train_rng, np.uint32), 3)
    return k[0], k[1], k[2]


def make_inputs(moments_nhwc, sample_rng, alphas_cumprod, num_train_timesteps=1000):
    """:16-43 -> (noisy_latents, noise, timesteps, latents), all NCHW float32 / int32."""
    moments = np.asarray(moments_nhwc, np.float32)
    mean, logvar = np.split(moments, 2, axis=-1)
    logvar = np.clip(logvar, np.float32(-30.0), np.float32(20.0))
    std = np.exp(np.float32(0.5) * logvar).astype(np.float32)
    latents = (mean + std * threefry.normal(sample_rng, mean.shape)).astype(np.float32)      # :17
    latents = (np.transpose(latents, (0, 3, 1, 2)) * VAE_SCALING).astype(np.float32)          # :19-20
    noise_rng, timestep_rng = threefry.split(sample_rng)                                      # :23
    noise = threefry.normal(noise_rng, latents.shape)                                         # :24
    timesteps = threefry.randint(timestep_rng, (latents.shape[0],), 0, num_train_timesteps)   # :27-32
    ac = np.asarray(alphas_cumprod, np.float32)[timesteps]
    sa = np.sqrt(ac).astype(np.float32).reshape(-1, 1, 1, 1)
    sb = np.sqrt(np.float32(1.0) - ac).astype(np.float32).reshape(-1, 1, 1, 1)
    noisy = (sa * latents + sb * noise).astype(np.float32)                                    # :36-41
    return noisy, noise, timesteps, latents


def mse_loss(eps_u, eps_c, noise, guidance_scale, train_cfg=True, weights=None):
    """:77-90 on torch tensors (differentiable)."""
    pred = eps_u + guidance_scale * (eps_c - eps_u) if train_cfg else eps_c
    loss = ((torch.as_tensor(noise).to(pred.dtype) - pred) ** 2).mean(dim=tuple(range(1, pred.dim())))
    if weights is None:
        return loss.mean(), loss
    w = torch.as_tensor(weights).to(pred.dtype)
    assert loss.numel() == w.numel()
    return (loss * w).sum(), loss


def train_loss(unet, batch_inputs, prompt_embeds, uncond_embeds, train_cfg, guidance_scale, weights=None):
    """U-Net part of compute_loss (:54-90): ``batch_inputs`` = (noisy, noise, timesteps) from ``make_inputs``."""
    noisy, noise, timesteps = batch_inputs[:3]
    x = torch.as_tensor(noisy)
    ts = torch.as_tensor(np.asarray(timesteps)).long()
    cond = unet(x, ts, torch.as_tensor(prompt_embeds))
    unc = unet(x, ts, torch.as_tensor(uncond_embeds)) if train_cfg else cond
    return mse_loss(unc, cond, noise, guidance_scale, train_cfg, weights)
