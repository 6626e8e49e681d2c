"""CPU restatement of the reference sampler ``FlaxStableDiffusionPipeline._generate``
(``ddpo/diffusers_patch/pipeline_flax_stable_diffusion.py:163-270``) and of the PPO
``train_step`` loss (``ddpo/training/policy_gradient.py:86-136``) on top of the oracle U-Net,
scheduler and PRNG.  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch

from . import ppo, scheduler as S, threefry
from .unet import UNetOracle


def generate(unet: UNetOracle, sched_cfg, sched_state, prompt_embeds, neg_prompt_embeds, rng, num_inference_steps,
             latent_hw, guidance_scale, eta, latents=None):
    """Returns (final_latents, latents [B,T,...], next_latents [B,T,...], log_probs [B,T], ts [B,T])."""
    b = prompt_embeds.shape[0]
    context = torch.cat([torch.as_tensor(neg_prompt_embeds), torch.as_tensor(prompt_embeds)])  # uncond first (:187)
    shape = (b, 4, latent_hw, latent_hw)
    rng = np.asarray(rng, np.uint32)
    if latents is None:
        rng, seed = threefry.split(rng)                                   # :196
        latents = threefry.normal(seed, shape)                            # :197
    state = S.set_timesteps(sched_cfg, sched_state, num_inference_steps)  # :243-247
    latents = (latents * state.init_noise_sigma).astype(np.float32)       # :250
    rng, carry = threefry.split(rng)                                      # :252
    lat_list, next_list, lp_list, t_list = [], [], [], []
    x = latents
    for step in range(num_inference_steps):                               # lax.scan body :204-241
        t = int(state.timesteps[step])
        xin = torch.from_numpy(np.concatenate([x, x]))
        with torch.no_grad():
            noise_pred = unet(xin, torch.full((2 * b,), t, dtype=torch.int64), context).float().numpy()
        eu, ec = noise_pred[:b], noise_pred[b:]
        eps = ppo.cfg_combine(eu, ec, guidance_scale)                     # :226-229
        carry, key = threefry.split(carry)                                # :232
        new_x, _, lp = S.step(sched_cfg, state, eps, t, x, key=key, eta=eta)
        lat_list.append(x), next_list.append(new_x), lp_list.append(lp), t_list.append(t)
        x = new_x
    lat = np.stack(lat_list, 1)
    nxt = np.stack(next_list, 1)
    lps = np.stack(lp_list, 1)
    ts = np.broadcast_to(np.asarray(t_list, np.int32), (b, num_inference_steps))
    return x, lat, nxt, lps, ts


def train_loss(unet: UNetOracle, sched_cfg, sched_state, batch, train_cfg, guidance_scale, eta, clip_range):
    """Differentiable (torch autograd) restatement of compute_loss (training/policy_gradient.py:86-136).
    batch: dict of numpy/torch arrays as the reference builds it (pipeline/policy_gradient.py:415-423).
    Returns (loss tensor, info dict, log_prob tensor)."""
    lat = torch.as_tensor(batch["latents"])
    ts = torch.as_tensor(np.asarray(batch["ts"])).long()
    cond = unet(lat, ts, torch.as_tensor(batch["prompt_embeds"]))
    if train_cfg:
        unc = unet(lat, ts, torch.as_tensor(batch["uncond_embeds"]))
        eps = unc + guidance_scale * (cond - unc)
    else:
        eps = cond
    dt = eps.dtype
    a_t, a_prev, sigma = S.coefficients(sched_cfg, sched_state, np.asarray(batch["ts"]), eta)
    tt = lambda v: torch.as_tensor(v, dtype=dt).view(-1, 1, 1, 1)
    a_t, a_prev, sigma = tt(a_t), tt(a_prev), tt(sigma)
    x = lat.to(dt)
    pred = getattr(sched_cfg, "prediction_type", "epsilon")               # scheduling_ddim_flax.py:303-321
    if pred == "epsilon":
        x0 = (x - torch.sqrt(1 - a_t) * eps) / torch.sqrt(a_t)
    elif pred == "sample":
        x0 = eps
    elif pred == "v_prediction":
        x0 = torch.sqrt(a_t) * x - torch.sqrt(1 - a_t) * eps
        eps = torch.sqrt(a_t) * eps + torch.sqrt(1 - a_t) * x
    else:
        raise ValueError(f"prediction_type given as {pred} must be one of `epsilon`, `sample`, or `v_prediction`")
    mean = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev - sigma ** 2) * eps
    sd = torch.clamp(sigma, min=1e-6)
    nxt = torch.as_tensor(batch["next_latents"]).to(dt)
    lp = -((nxt - mean) ** 2) / (2 * sd ** 2) - torch.log(sd) - float(np.log(np.sqrt(2 * np.pi)))
    log_prob = lp.mean(dim=(1, 2, 3))
    adv = torch.clamp(torch.as_tensor(batch["advantages"]).to(dt), -ppo.ADV_CLIP_MAX, ppo.ADV_CLIP_MAX)
    ratio = torch.exp(log_prob - torch.as_tensor(batch["log_probs"]).to(dt))
    unclipped = -adv * ratio
    clipped = -adv * torch.clamp(ratio, 1.0 - clip_range, 1.0 + clip_range)
    loss = torch.maximum(unclipped, clipped).mean()
    info = {"approx_kl": 0.5 * ((log_prob - torch.as_tensor(batch["log_probs"]).to(dt)) ** 2).mean(),
            "clipfrac": (torch.abs(ratio - 1.0) > clip_range).to(dt).mean(), "loss": loss}
    return loss, info, log_prob
