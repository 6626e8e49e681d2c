"""The public API's host logic end to end on the CPU ops emulator: ``pipeline(...)`` (key lineage, trajectory buffers,
per-step scheduler calls) against the oracle sampler, then ``train_step(...)`` on the sampled trajectory (batch
assembly, CFG ordering, PPO gradient routing, accumulate / update bookkeeping).  The CUDA kernels are checked on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))


@pytest.fixture()
def emulated(monkeypatch):
    import _cpu_ops_emulator as E
    from ddpo_b200 import unet as U
    from ddpo_b200.diffusers_patch import pipeline_stable_diffusion as P, scheduling_ddim as SD
    from ddpo_b200.training import policy_gradient as PG
    for mod in (U, P, SD, PG):
        monkeypatch.setattr(mod, "ops", E)
    monkeypatch.setattr(U, "Arena", E.CpuArena)
    monkeypatch.setattr(PG, "USE_CUDA_GRAPH", False)
    PG._GRAPHS.clear()
    return U, P, SD, PG


def test_pipeline_and_train_step_host_logic(emulated):
    U, P, SD, PG = emulated
    from ddpo_b200 import unet_spec
    from oracle import pipeline as OP, scheduler as OS
    from oracle.unet import UNetOracle
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator().manual_seed(1)
    b, T = 2, 3
    emb = torch.randn(b, cfg.ctx_len, cfg.cross_attention_dim, generator=g)
    neg = torch.randn(1, cfg.ctx_len, cfg.cross_attention_dim, generator=g).expand(b, -1, -1).contiguous()
    net = U.UNet(cfg, flat, device="cpu")
    sched = SD.DDIMScheduler(1000, 0.00085, 0.012, "scaled_linear", None, False, 1, "epsilon", device="cpu")
    pipe = P.StableDiffusionPipeline(net, sched, use_cuda_graph=False)
    state = sched.create_state()
    px = cfg.sample_size * 8
    final, lat, nxt, lps, ts = pipe(emb, neg, {"unet": net.params, "scheduler": state}, (0, 7), T, px, px, 5.0, 1.0)
    onet = UNetOracle(cfg, unet_spec.views(flat, cfg))
    rf, rlat, rnxt, rlps, rts = OP.generate(onet, OS.SD_CONFIG, OS.create_state(OS.SD_CONFIG), emb, neg,
                                            np.array((0, 7), np.uint32), T, cfg.sample_size, 5.0, 1.0)
    assert tuple(lat.shape) == rlat.shape and tuple(lps.shape) == rlps.shape
    np.testing.assert_array_equal(ts.numpy(), rts)
    np.testing.assert_allclose(lat[:, 0].numpy(), rlat[:, 0], atol=3e-6, rtol=0)          # x_T: same threefry draw
    np.testing.assert_allclose(lps.numpy(), rlps, rtol=1e-3)
    rel1 = np.linalg.norm(nxt[:, 0].numpy() - rnxt[:, 0]) / np.linalg.norm(rnxt[:, 0])
    assert rel1 < 2e-2, rel1
    assert torch.equal(lat[:, 1:], nxt[:, :-1])
    # ---- PPO update on the sampled trajectory: unchanged policy -> ratio == 1 exactly
    st = sched.set_timesteps(state, T)
    tstate = PG.AccumulatingTrainState(apply_fn=net, tx=PG.AdamWConfig(learning_rate=1e-3))
    batch = {"latents": lat[:, 1].contiguous(), "next_latents": nxt[:, 1].contiguous(), "ts": ts[:, 1].contiguous(),
             "log_probs": lps[:, 1].contiguous(), "advantages": torch.tensor([1.5, -0.7]),
             "prompt_embeds": emb, "uncond_embeds": neg}
    p0 = net.params.clone()
    tstate, info = PG.train_step(tstate, batch, st, sched, True, 5.0, 1.0, 1e-4, False)
    assert info["approx_kl"].item() == 0.0 and info["clipfrac"].item() == 0.0
    assert abs(info["loss"].item() - (-(1.5 - 0.7) / 2)) < 1e-6
    assert tstate.n_acc == 1 and tstate.step == 0 and torch.equal(net.params, p0) and float(net.grads.abs().max()) > 0
    g1 = net.grads.clone()
    # a macro batch of two timesteps (extension) accumulates both micro-batch gradients; then the update fires
    both = {k: (torch.cat([batch[k], batch[k]]) if k not in ("latents", "next_latents", "ts", "log_probs") else
                torch.cat([batch[k], {"latents": lat[:, 2], "next_latents": nxt[:, 2], "ts": ts[:, 2],
                                      "log_probs": lps[:, 2]}[k].contiguous()])) for k in batch}
    tstate, info2 = PG.train_step(tstate, both, st, sched, True, 5.0, 1.0, 1e-4, True, micro_batch_size=2)
    assert info2["approx_kl"].item() == 0.0
    assert tstate.step == 1 and tstate.n_acc == 0 and float(net.grads.abs().max()) == 0.0
    assert not torch.equal(net.params, p0)
    assert float(tstate.last_grad_norm) > 0
    assert float((g1 != 0).float().mean()) > 0.5
