"""NumPy float32 restatement of the reference DDIM scheduler with log-prob.

Follows ``ddpo/diffusers_patch/scheduling_ddim_flax.py`` (reference):
  * ``create_state``   :144-170  (+ diffusers 0.12.1 ``CommonSchedulerState.create``)
  * ``set_timesteps``  :189-211
  * ``_get_variance``  :213-227
  * ``step``           :229-361  (sample mode with ``key`` / score mode with ``prev_sample``)
plus the analytic d log_prob / d model_output used to check the CUDA backward.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np

from . import threefry

f32 = np.float32


@dataclass
class SchedulerConfig:
    num_train_timesteps: int = 1000
    beta_start: float = 0.0001
    beta_end: float = 0.02
    beta_schedule: str = "linear"
    set_alpha_to_one: bool = True
    steps_offset: int = 0
    prediction_type: str = "epsilon"


# Stable Diffusion's scheduler_config.json (v1.x and 2-base share these values)
SD_CONFIG = SchedulerConfig(1000, 0.00085, 0.012, "scaled_linear", False, 1, "epsilon")


@dataclass
class SchedulerState:
    alphas_cumprod: np.ndarray
    final_alpha_cumprod: np.float32
    init_noise_sigma: np.float32
    timesteps: np.ndarray
    num_inference_steps: Optional[int] = None


def create_state(cfg: SchedulerConfig) -> SchedulerState:
    n = cfg.num_train_timesteps
    if cfg.beta_schedule == "linear":
        betas = np.linspace(cfg.beta_start, cfg.beta_end, n, dtype=f32)
    elif cfg.beta_schedule == "scaled_linear":
        betas = np.linspace(cfg.beta_start ** 0.5, cfg.beta_end ** 0.5, n, dtype=f32) ** 2
    else:
        raise NotImplementedError(cfg.beta_schedule)
    alphas = (f32(1.0) - betas).astype(f32)
    ac = np.cumprod(alphas, dtype=f32)
    final = f32(1.0) if cfg.set_alpha_to_one else ac[0]
    ts = np.arange(0, n)[::-1].copy()
    return SchedulerState(ac, final, f32(1.0), ts, None)


def set_timesteps(cfg, state, num_inference_steps):
    ratio = cfg.num_train_timesteps // num_inference_steps
    ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1] + cfg.steps_offset
    return replace(state, num_inference_steps=num_inference_steps, timesteps=ts.astype(np.int64))


def _bcast(v, ndim):
    v = np.asarray(v, f32)
    return v.reshape(v.shape + (1,) * (ndim - v.ndim))


def coefficients(cfg, state, timestep, eta):
    """Per-sample scalars of one step: (alpha_t, alpha_prev, sigma)."""
    t = np.asarray(timestep)
    prev_t = t - cfg.num_train_timesteps // state.num_inference_steps
    a_t = state.alphas_cumprod[t]
    a_prev = np.where(prev_t >= 0, state.alphas_cumprod[np.maximum(prev_t, 0)],
                      state.final_alpha_cumprod).astype(f32)
    b_t = (f32(1) - a_t).astype(f32)
    b_prev = (f32(1) - a_prev).astype(f32)
    var = ((b_prev / b_t) * (f32(1) - a_t / a_prev)).astype(f32)
    sigma = (f32(eta) * np.sqrt(var)).astype(f32)
    return a_t.astype(f32), a_prev, sigma


def step(cfg, state, model_output, timestep, sample, key=None, prev_sample=None, eta=0.0,
         return_mean=False):
    if state.num_inference_steps is None:
        raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps'")
    if prev_sample is not None and key is not None:
        raise ValueError("Cannot pass both key and prev_sample.")
    x = np.asarray(sample, f32)
    eps = np.asarray(model_output, f32)
    a_t, a_prev, sigma = coefficients(cfg, state, timestep, eta)
    a_t, a_prev, sigma = (_bcast(v, x.ndim) for v in (a_t, a_prev, sigma))
    b_t = f32(1) - a_t
    if cfg.prediction_type == "epsilon":                                   # :303-306
        x0 = (x - np.sqrt(b_t) * eps) / np.sqrt(a_t)
    elif cfg.prediction_type == "sample":                                  # :307-308 (model_output stays as is)
        x0 = eps
    elif cfg.prediction_type == "v_prediction":                            # :309-316
        x0 = np.sqrt(a_t) * x - np.sqrt(b_t) * eps
        eps = np.sqrt(a_t) * eps + np.sqrt(b_t) * x
    else:
        raise ValueError(f"prediction_type given as {cfg.prediction_type} must be one of `epsilon`, `sample`, or"
                         " `v_prediction`")
    direction = np.sqrt(f32(1) - a_prev - sigma ** 2) * eps
    mean = (np.sqrt(a_prev) * x0 + direction).astype(f32)
    if prev_sample is None:
        noise = threefry.normal(key, x.shape)
        prev_sample = (mean + sigma * noise).astype(f32)
    prev_sample = np.asarray(prev_sample, f32)
    sd = np.maximum(sigma, f32(1e-6)).astype(f32)
    lp = (-((prev_sample - mean) ** 2) / (f32(2) * sd ** 2) - np.log(sd)
          - np.log(np.sqrt(f32(2 * np.pi), dtype=f32))).astype(f32)
    lp = np.broadcast_to(lp, x.shape)
    log_prob = lp.reshape(x.shape[0], -1).mean(axis=1, dtype=f32) if x.ndim > 1 else lp.mean(dtype=f32)
    if return_mean:
        return prev_sample, state, log_prob, mean
    return prev_sample, state, log_prob


def logprob_grad_eps(cfg, state, model_output, timestep, sample, prev_sample, eta, dlogp):
    """d(sum_b dlogp[b] * log_prob[b]) / d model_output  (float64 analytic)."""
    x = np.asarray(sample, np.float64)
    eps = np.asarray(model_output, np.float64)
    a_t, a_prev, sigma = coefficients(cfg, state, timestep, eta)
    a_t, a_prev, sigma = (_bcast(v, x.ndim).astype(np.float64) for v in (a_t, a_prev, sigma))
    d = np.sqrt(1 - a_prev - sigma ** 2)
    if cfg.prediction_type == "epsilon":
        c_eps = d - np.sqrt(a_prev) * np.sqrt(1 - a_t) / np.sqrt(a_t)
        mean = np.sqrt(a_prev) * (x - np.sqrt(1 - a_t) * eps) / np.sqrt(a_t) + d * eps
    elif cfg.prediction_type == "sample":
        c_eps = np.sqrt(a_prev) + d
        mean = c_eps * eps
    else:
        c_eps = d * np.sqrt(a_t) - np.sqrt(a_prev) * np.sqrt(1 - a_t)
        mean = np.sqrt(a_prev) * (np.sqrt(a_t) * x - np.sqrt(1 - a_t) * eps) + d * (np.sqrt(a_t) * eps + np.sqrt(1 - a_t) * x)
    sd = np.maximum(sigma, 1e-6)
    n = x[0].size
    g = (np.asarray(prev_sample, np.float64) - mean) / sd ** 2 * c_eps / n
    return g * _bcast(np.asarray(dlogp, np.float64), x.ndim)
