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
    assert float(net.grads.abs().max()) > 0.0


def _ddim_mean(eps, lat, coef):
    """epsilon-prediction DDIM mean (scheduling_ddim_flax.py:303-343) from already-combined eps; coef = (a_t, a_prev, sigma)"""
    a_t, a_prev, sigma = [torch.as_tensor(np.asarray(c), dtype=torch.float32).view(-1, 1, 1, 1) for c in coef]
    x0 = (lat - torch.sqrt(1 - a_t) * eps) / torch.sqrt(a_t)
    return torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev - sigma ** 2) * eps


def test_ppo_gradient_through_the_loss_matches_oracle():
    """The whole PPO path -- U-Net (cond + uncond), CFG, score-mode log-prob, clipped surrogate, backward -- against the
    oracle's autograd gradient (`_oracle_grads`).

    d loss / d eps ~ adv * (x_prev - mean(eps)) / sigma^2.  On the CUDA side x_prev - mean is the sampler's sigma * z
    exactly (the mean is reproduced bit for bit); handing the fp32 oracle the bf16 sampler's x_prev would make ITS
    x_prev - mean = sigma * z + (mean_gpu - mean_oracle), and at t = 334 -> 1 (sigma = 0.03) that difference is as large as
    the noise term: the two sides would differentiate different points of the loss (measured that way: cosine 0.88).  So the
    oracle gets the SAME noise term around its own mean -- x_prev' = mean_oracle + (x_prev - mean_gpu) -- and its own
    log-prob as the old one (ratio == 1 on both sides): what is compared is then the loss -> eps -> parameters chain itself."""
    from ddpo_b200 import unet_spec
    from ddpo_b200.training import policy_gradient as pg
    from oracle import scheduler as OS
    from oracle.unet import UNetOracle
    pg.USE_CUDA_GRAPH = False
    pg._GRAPHS.clear()
    cfg, flat, emb, neg, net, sched, st, out = _sample()
    state = pg.AccumulatingTrainState(apply_fn=net)
    adv = [1.5, -0.7]
    batch = _batch(out, emb, neg, 1, adv)
    state, info = pg.train_step(state, batch, st, sched, True, 5.0, 1.0, 1e-4, False)
    torch.cuda.synchronize()
    assert info["approx_kl"].item() == 0.0
    g_gpu = net.grads.cpu().clone()
    # the sampler's noise term sigma * z = x_prev - mean_gpu (fp32 formulas on the CUDA eps: equal to the kernel's to rounding)
    lat, ts = batch["latents"], batch["ts"]
    net.prepare_context(torch.cat([neg, emb]).cuda())
    e2 = net.forward(torch.cat([lat, lat]), torch.cat([ts, ts])).float().cpu()
    b = lat.shape[0]
    ost = OS.set_timesteps(OS.SD_CONFIG, OS.create_state(OS.SD_CONFIG), 3)
    coef = OS.coefficients(OS.SD_CONFIG, ost, ts.cpu().numpy(), 1.0)
    mean_gpu = _ddim_mean(e2[:b] + 5.0 * (e2[b:] - e2[:b]), lat.cpu(), coef)
    noise = batch["next_latents"].cpu() - mean_gpu
    with torch.no_grad():
        onet = UNetOracle(cfg, unet_spec.views(flat, cfg))
        tsl = ts.cpu().long()
        oc, ou = onet(lat.cpu(), tsl, emb), onet(lat.cpu(), tsl, neg)
        mean_ref = _ddim_mean(ou + 5.0 * (oc - ou), lat.cpu(), coef)
    raw = ((mean_gpu - mean_ref).norm() / noise.norm()).item()
    obatch = dict(batch)
    obatch["next_latents"] = mean_ref + noise
    g_ref, rinfo, _ = _oracle_grads(cfg, flat, obatch, True, 1e-4)
    assert abs(rinfo["approx_kl"]) < 1e-12 and abs(rinfo["loss"] - info["loss"].item()) < 1e-5
    rel = ((g_gpu - g_ref).norm() / g_ref.norm()).item()
    cos = (torch.dot(g_gpu, g_ref) / (g_gpu.norm() * g_ref.norm())).item()
    ratio = (g_gpu.norm() / g_ref.norm()).item()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ppo_grad_parity_TINY.txt", "w") as f:
        f.write(f"PPO gradient through the loss (TINY, same noise term): rel-L2 {rel:.3e} cosine {cos:.6f} norm ratio "
                f"{ratio:.4f}; |mean_gpu - mean_oracle| / |sigma z| = {raw:.3f}\n")
    assert cos > 0.99 and abs(ratio - 1) < 0.05 and rel < 0.15, (rel, cos, ratio)


def _backward_vs_oracle(cfg_name, batch, tag):
    """U-Net backward in isolation: same upstream gradient d_eps on both sides.  (Comparing through the PPO loss
    is ill-conditioned: d_eps ~ (x_prev - mean(eps)) and the fp32-vs-bf16 difference of eps is amplified by the CFG
    combine (x5/-4) and by c_eps/sigma, so the two losses are evaluated at visibly different points.)"""
    from ddpo_b200 import unet_spec
    from ddpo_b200.unet import UNet
    from oracle.unet import UNetOracle
    cfg = getattr(unet_spec, cfg_name)
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator(device="cpu").manual_seed(11)
    s = cfg.sample_size
    lat = torch.randn(batch, 4, s, s, generator=g)
    ctx = torch.randn(batch, cfg.ctx_len, cfg.cross_attention_dim, generator=g)
    ts = torch.tensor([981, 441, 21, 1][:batch], dtype=torch.int32)
    d_eps = torch.randn(batch, 4, s, s, generator=g) / (4 * s * s)
    net = UNet(cfg, flat, "cuda")
    net.enable_training()
    net.prepare_context(ctx.cuda())
    tape = []
    net._dbg = {}
    eps = net.forward(lat.cuda(), ts.cuda(), tape=tape)
    net.backward(tape, d_eps.cuda())
    torch.cuda.synchronize()
    fp = flat.clone().requires_grad_(True)
    taps = {}

    def tap(n, t):
        if t.requires_grad:
            t.retain_grad()
        taps[n] = t
    ref = UNetOracle(cfg, unet_spec.views(fp, cfg), tap=tap)(lat, ts, ctx)
    ref.backward(d_eps)
    table, _ = unet_spec.param_offsets(cfg)
    gg = net.grads.cpu()
    lines, worst, worst_name = [], 0.0, None
    for name in reversed([n for n in taps if n in net._dbg and taps[n].grad is not None]):
        r = taps[name].grad
        a = net._dbg[name].cpu().view(r.shape)
        lines.append(f"act {name}\t{((a - r).norm() / (r.norm() + 1e-30)).item():.3e}")
    for name, (off, shape) in table.items():
        n = int(np.prod(shape))
        a, r = gg[off:off + n], fp.grad[off:off + n]
        e = ((a - r).norm() / (r.norm() + 1e-30)).item()
        lines.append(f"{name}\t{e:.3e}\t{r.norm().item():.3e}")
        if r.norm().item() > 1e-3 * fp.grad.norm().item() / np.sqrt(len(table)) and e > worst:
            worst, worst_name = e, name
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/grad_parity_{tag}.txt", "w") as f:
        f.write("\n".join(lines))
    tot = ((gg - fp.grad).norm() / fp.grad.norm()).item()
    print(f"{tag}: total grad rel err {tot:.3e}; worst tensor {worst_name} {worst:.3e}")
    return tot, worst, worst_name


@pytest.mark.parametrize("cfg_name,batch", [("TINY", 2), ("TINY", 3), ("SMALL", 2)])
def test_unet_backward_matches_oracle(cfg_name, batch):
    tot, worst, worst_name = _backward_vs_oracle(cfg_name, batch, f"{cfg_name}_b{batch}")
    assert tot < 4e-2, f"total gradient relative error {tot}"
    assert worst < 0.2, f"worst tensor {worst_name}: {worst}"


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
    np.testing.assert_allclose(net.params.cpu().numpy(), ost.params, rtol=0, atol=2e-5)
    assert np.abs(net.params.cpu().numpy() - ost.params).mean() < 2e-8
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
    assert np.isfinite(info["loss"].item()) and float(net.grads.abs().max()) > 0.0
    assert torch.isfinite(net.grads).all()


def test_macro_batched_train_step_equals_consecutive_calls():
    """Stacking the same 2 samples at 3 timesteps into one pass (micro_batch_size=2) accumulates the same gradient,
    n_acc and mean info as three reference-sized calls."""
    from ddpo_b200.training import policy_gradient as pg
    from ddpo_b200.unet import UNet
    pg.USE_CUDA_GRAPH = False
    pg._GRAPHS.clear()
    cfg, flat, emb, neg, net, sched, st, out = _sample()
    advs = [[1.0, -1.0], [0.3, 2.0], [-0.5, 0.5]]
    state = pg.AccumulatingTrainState(apply_fn=net)
    infos = []
    for j in range(3):
        _, info = pg.train_step(state, _batch(out, emb, neg, j, advs[j]), st, sched, True, 5.0, 1.0, 1e-4, False)
        infos.append(info["loss"].item())
    g_seq = net.grads.clone()
    assert state.n_acc == 3
    net2 = UNet(cfg, flat, "cuda")
    state2 = pg.AccumulatingTrainState(apply_fn=net2)
    bs = [_batch(out, emb, neg, j, advs[j]) for j in range(3)]
    big = {k: torch.cat([b[k] for b in bs]) for k in bs[0]}
    _, info = pg.train_step(state2, big, st, sched, True, 5.0, 1.0, 1e-4, False, micro_batch_size=2)
    torch.cuda.synchronize()
    assert state2.n_acc == 3
    assert info["approx_kl"].item() == 0.0
    np.testing.assert_allclose(info["loss"].item(), np.mean(infos), rtol=1e-5, atol=1e-6)
    rel = ((net2.grads - g_seq).norm() / g_seq.norm()).item()
    assert rel < 1e-3, f"macro-batched gradient differs from consecutive calls by {rel}"
