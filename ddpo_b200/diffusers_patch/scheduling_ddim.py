"""DDIM scheduler with per-step Gaussian log-prob -- drop-in mirror of the reference's
``ddpo/diffusers_patch/scheduling_ddim_flax.py`` (``FlaxDDIMScheduler``: ``create_state`` :144-170,
``set_timesteps`` :189-211, ``step`` :229-361).  Same names, argument meaning and error behaviour;
the arithmetic of ``step`` runs in the fused CUDA kernels of ``csrc/ddim.cu``.
"""
from dataclasses import dataclass, replace
from types import SimpleNamespace
from typing import Optional, Tuple

import numpy as np
import torch

from .. import ops


@dataclass
class CommonSchedulerState:
    alphas_cumprod: torch.Tensor  # fp32 [num_train_timesteps] on the device


@dataclass
class DDIMSchedulerState:
    common: CommonSchedulerState
    final_alpha_cumprod: float
    init_noise_sigma: float
    timesteps: np.ndarray
    num_inference_steps: Optional[int] = None

    def replace(self, **kw):
        return replace(self, **kw)


class DDIMScheduler:
    has_state = True

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, set_alpha_to_one: bool = True,
                 steps_offset: int = 0, prediction_type: str = "epsilon", dtype=torch.float32, device="cuda",
                 **kwargs):
        if "predict_epsilon" in kwargs:  # deprecated alias handled by the reference (:130-140)
            pe = kwargs.pop("predict_epsilon")
            if pe is not None:
                prediction_type = "epsilon" if pe else "sample"
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, trained_betas=trained_betas,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type=prediction_type)
        self.dtype = dtype
        self.device = device
        self._ws = {}
        self._one = None

    # --- reference: create_state (:144-170) + diffusers CommonSchedulerState.create
    def create_state(self, common: Optional[CommonSchedulerState] = None) -> DDIMSchedulerState:
        c = self.config
        if common is None:
            if c.trained_betas is not None:
                betas = np.asarray(c.trained_betas, np.float32)
            elif c.beta_schedule == "linear":
                betas = np.linspace(c.beta_start, c.beta_end, c.num_train_timesteps, dtype=np.float32)
            elif c.beta_schedule == "scaled_linear":
                betas = np.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps,
                                    dtype=np.float32) ** 2
            else:
                raise NotImplementedError(f"beta_schedule {c.beta_schedule} is not implemented")
            ac = np.cumprod((np.float32(1.0) - betas).astype(np.float32), dtype=np.float32)
            common = CommonSchedulerState(torch.from_numpy(ac).to(self.device))
        ac0 = float(common.alphas_cumprod[0].item())
        final = 1.0 if c.set_alpha_to_one else ac0
        ts = np.arange(0, c.num_train_timesteps).round()[::-1].copy()
        return DDIMSchedulerState(common=common, final_alpha_cumprod=final, init_noise_sigma=1.0, timesteps=ts)

    def scale_model_input(self, state, sample, timestep=None):
        return sample

    # --- reference: set_timesteps (:189-211)
    def set_timesteps(self, state: DDIMSchedulerState, num_inference_steps: int, shape: Tuple = ()):
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1] + self.config.steps_offset
        return state.replace(num_inference_steps=num_inference_steps, timesteps=ts.astype(np.int64).copy())

    def _workspace(self, batch, device):
        k = (batch, str(device))
        if k not in self._ws:
            self._ws[k] = ops.ddim_workspace(batch, device)
        return self._ws[k]

    # --- reference: step (:229-361)
    def step(self, state: DDIMSchedulerState, model_output: torch.Tensor, timestep, sample: torch.Tensor, key=None,
             prev_sample: Optional[torch.Tensor] = None, eta: float = 0.0):
        if state.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        pred = self.config.prediction_type
        ops.prediction_code(pred)  # ValueError for anything but epsilon / sample / v_prediction (reference :317-321)
        if prev_sample is not None and key is not None:
            raise ValueError("Cannot pass both key and prev_sample. Please make sure that either `key` or"
                             " `prev_sample` stays `None`.")
        if prev_sample is None and key is None:
            raise ValueError("Either `key` or `prev_sample` is required.")
        dev = sample.device
        b = sample.shape[0]
        eps = model_output.contiguous().float()
        x = sample.contiguous().float()
        if torch.is_tensor(timestep):
            ts = timestep.to(dev, torch.int32).reshape(-1).contiguous()
        else:
            ts = torch.as_tensor(np.asarray(timestep, np.int32).reshape(-1), device=dev)
        ratio = self.config.num_train_timesteps // state.num_inference_steps
        ws = self._workspace(b, dev)
        logp = torch.empty(b, dtype=torch.float32, device=dev)
        ac = state.common.alphas_cumprod
        # model_output is already guidance-combined: eps_u == eps_c with guidance 0 -> eps exactly
        if prev_sample is None:
            kd = key if torch.is_tensor(key) else ops.key_tensor([tuple(int(v) for v in key)], dev)
            prev = torch.empty_like(x)
            ops.ddim_step_sample(eps, eps, x, ac, ts, state.final_alpha_cumprod, ratio, 0.0, float(eta), kd, prev,
                                 logp, ws, pred=pred)
        else:
            prev = prev_sample.contiguous().float()
            ops.ddim_logprob_fwd(eps, eps, x, prev, ac, ts, state.final_alpha_cumprod, ratio, 0.0, float(eta), logp, ws,
                                 pred=pred)
        return prev, state, logp

    def __len__(self):
        return self.config.num_train_timesteps


FlaxDDIMScheduler = DDIMScheduler  # the reference's class name
