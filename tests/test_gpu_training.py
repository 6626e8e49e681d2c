"""PPO train step (CUDA fwd+bwd+optimizer) vs the oracle (torch-CPU autograd + NumPy AdamW)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sample(cfg_name="TINY", batch=2, T=3, seed=0):
    from ddpo_b200 import unet_spec
    from ddpo_b200.diffusers_patch import DDIMScheduler, StableDiffusionPipeline
    from ddpo_b200.unet import UNet
    cfg = getattr(unet_spec, cfg_name)
    flat = unet_spec.init_flat_params(cfg, seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    emb = torch.randn(batch, cfg.ctx_len, cfg.cross_attention_dim, generator=g)
    neg = torch.randn(1, cfg.ctx_len, cfg.cross_attention_dim, generator=g).expand(batch, -1, -1).contiguous()
    net = UNet(cfg, flat, "cuda")
    sched = DDIMScheduler(1000, 0.00085, 0.012, "scaled_linear", None, False, 1, "epsilon")
    pipe = StableDiffusionPipeline(net, sched, vae_scale_factor=8, use_cuda_graph=False)
    state = sched.create_state()
    px = cfg.sample_size * 8
    out = pipe(emb.cuda(), neg.cuda(), {"unet": net.params, "scheduler": state}, (3, 4), T, px, px, 5.0, 1.0)
    return cfg, flat, emb, neg, net, sched, sched.set_timesteps(state, T), [o.clone() for o in out]


def _batch(out, emb, neg, j, adv):
    final, lat, nxt, lps, ts = out
    return {"latents": lat[:, j].contiguous(), "next_latents": nxt[:, j].contiguous(), "ts": ts[:, j].contiguous(),
            "log_probs": lps[:, j].contiguous(), "advantages": torch.tensor(adv, device="cuda"),
            "prompt_embeds": emb.cuda(), "uncond_embeds": neg.cuda()}


def _oracle_grads(cfg, flat, batch, train_cfg, clip, self_consistent_old=True):
    from ddpo_b200 import unet_spec
    from oracle import pipeline as OP, scheduler as OS
    from oracle.unet import UNetOracle
    fp = flat.clone().requires_grad_(True)
    onet = UNetOracle(cfg, unet_spec.views(fp, cfg))
    ost = OS.set_timesteps(OS.SD_CONFIG, OS.create_state(OS.SD_CONFIG), 3)
    nb = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in batch.items()}
    if self_consistent_old:
        # the CUDA path reproduces ITS OWN sampled log-prob bit-exactly (ratio == 1).  The fp32 oracle differs from
        # the bf16 sampler by ~1e-4 relative, which is the size of the clip range: give the oracle its own
        # log-prob as the "old" one so that both sides differentiate the same (unclipped, ratio == 1) branch.
        with torch.no_grad():
            _, _, lp0 = OP.train_loss(onet, OS.SD_CONFIG, ost, nb, train_cfg, 5.0, 1.0, clip)
        nb["log_probs"] = lp0.numpy()
    loss, info, lp = OP.train_loss(onet, OS.SD_CONFIG, ost, nb, train_cfg, 5.0, 1.0, clip)
    loss.backward()
    return fp.grad, {k: float(v) for k, v in info.items()}, lp.detach().numpy()


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_first_pass_ratio_is_one_and_grads_match_oracle(use_graph):
    from ddpo_b200 import unet_spec
    from ddpo_b200.training import policy_gradient as pg
    pg.USE_CUDA_GRAPH = use_graph
    pg._GRAPHS.clear()
    cfg, flat, emb, neg, net, sched, st, out = _sample()
    state = pg.AccumulatingTrainState(apply_fn=net)
    batch = _batch(out, emb, neg, 1, [1.5, -0.7])
    state, info = pg.train_step(state, batch, st, sched, True, 5.0, 1.0, 1e-4, False)
    torch.cuda.synchronize()
    # unchanged policy: log-prob reproduced bit-exactly -> ratio == 1
    assert info["approx_kl"].item() == 0.0 and info["clipfrac"].item() == 0.0
    assert abs(info["loss"].item() - (-(1.5 - 0.7) / 2)) < 1e-6
    gref, rinfo, _ = _oracle_grads(cfg, flat, batch, True, 1e-4)
    table, _ = unet_spec.param_offsets(cfg)
    g = net.grads.cpu()
    worst, lines = 0.0, []
    for name, (off, shape) in table.items():
        n = int(np.prod(shape))
        a, r = g[off:off + n], gref[off:off + n]
        e = ((a - r).norm() / (r.norm() + 1e-20)).item()
        lines.append(f"{name}\t{e:.3e}\t{r.norm().item():.3e}")
        if r.norm().item() > 1e-7:
            worst = max(worst, e)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/grad_parity_TINY_graph{int(use_graph)}.txt", "w") as f:
        f.write("\n".join(lines))
    tot = ((g - gref).norm() / gref.norm()).item()
    print(f"grad rel err total {tot:.3e} worst tensor {worst:.3e}")
    assert tot < 5e-2, f"total gradient relative error {tot}"
    assert worst < 0.25, f"worst per-tensor gradient relative error {worst}"


def test_train_step_accumulate_update_matches_oracle_optimizer():
    from ddpo_b200.training import policy_gradient as pg
    from oracle import optim as OO
    pg.USE_CUDA_GRAPH = True
    pg._GRAPHS.clear()
    cfg, flat, emb, neg, net, sched, st, out = _sample()
    state = pg.AccumulatingTrainState(apply_fn=net, tx=pg.AdamWConfig(learning_rate=1e-3))
    p0 = net.params.clone()
    pg.train_step(state, _batch(out, emb, neg, 0, [1.0, -1.0]), st, sched, True, 5.0, 1.0, 1e-4, False)
    g1 = net.grads.clone()
    pg.train_step(state, _batch(out, emb, neg, 2, [0.3, 2.0]), st, sched, True, 5.0, 1.0, 1e-4, False)
    gsum = net.grads.clone()
    assert state.n_acc == 2 and not torch.equal(g1, gsum)
    # third call triggers the update: g~ = (acc + g3) / 3
    state, info = pg.train_step(state, _batch(out, emb, neg, 1, [-0.5, 0.5]), st, sched, True, 5.0, 1.0, 1e-4, True)
    torch.cuda.synchronize()
    assert state.n_acc == 0 and state.step == 1 and float(net.grads.abs().max()) == 0.0
    # replay the same accumulated gradient through the oracle optimizer
    pg._GRAPHS.clear()
    net2_grads = None
    ost = OO.AccumulatingTrainState(p0.cpu().numpy(), lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4, max_norm=1.0)
    # recompute the three micro-gradients on a fresh copy of the network
    from ddpo_b200.unet import UNet
    netb = UNet(cfg, p0, "cuda")
    sb = pg.AccumulatingTrainState(apply_fn=netb)
    gs = []
    for j, adv in ((0, [1.0, -1.0]), (2, [0.3, 2.0]), (1, [-0.5, 0.5])):
        netb.grads.zero_()
        pg.train_step(sb, _batch(out, emb, neg, j, adv), st, sched, True, 5.0, 1.0, 1e-4, False)
        gs.append(netb.grads.cpu().numpy().copy())
    ost.apply_gradients(gs[0], False)
    ost.apply_gradients(gs[1], False)
    gn = ost.apply_gradients(gs[2], True)
    np.testing.assert_allclose(state.last_grad_norm.item(), gn, rtol=1e-4)
    np.testing.assert_allclose(net.params.cpu().numpy(), ost.params, rtol=0, atol=5e-6)
    mu = state.opt_state["mu"].float().cpu().numpy()
    np.testing.assert_allclose(mu, ost.opt.mu, rtol=1e-2, atol=1e-9)
    # the policy changed: next pass has ratio != 1
    state, info = pg.train_step(state, _batch(out, emb, neg, 1, [-0.5, 0.5]), st, sched, True, 5.0, 1.0, 1e-4, False)
    assert info["approx_kl"].item() > 0.0


def test_train_step_clipped_branch_and_no_cfg():
    """ratio far outside the clip range with the sign that selects the clipped branch -> zero gradient for that
    sample; train_cfg=False path (reference config 'train' default)."""
    from ddpo_b200.training import policy_gradient as pg
    pg.USE_CUDA_GRAPH = False
    pg._GRAPHS.clear()
    cfg, flat, emb, neg, net, sched, st, out = _sample()
    state = pg.AccumulatingTrainState(apply_fn=net)
    batch = _batch(out, emb, neg, 1, [1.0, 1.0])
    batch["log_probs"] = batch["log_probs"] - 0.01  # ratio = e^0.01 > 1 + clip, adv > 0 -> clipped, grad 0
    state, info = pg.train_step(state, batch, st, sched, True, 5.0, 1.0, 1e-4, False)
    torch.cuda.synchronize()
    assert info["clipfrac"].item() == 1.0
    assert float(net.grads.abs().max()) == 0.0
    # without CFG the log-prob differs from the (CFG-)sampled one: disable clipping so that both sides take the
    # unclipped branch  -A * ratio  whatever the ratio is
    batch = _batch(out, emb, neg, 1, [1.0, -2.0])
    state, info = pg.train_step(state, batch, st, sched, False, 5.0, 1.0, 1e9, False)
    torch.cuda.synchronize()
    gref, rinfo, rlp = _oracle_grads(cfg, flat, batch, False, 1e9, self_consistent_old=False)
    np.testing.assert_allclose(info["loss"].item(), rinfo["loss"], rtol=2e-2, atol=1e-3)
    tot = ((net.grads.cpu() - gref).norm() / gref.norm()).item()
    assert tot < 8e-2, f"no-cfg gradient relative error {tot}"
