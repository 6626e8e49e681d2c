"""RWR train step (ddpo_b200/training/diffusion.py + csrc/rwr.cu) vs the oracle restatement of the reference's
``ddpo/training/diffusion.py:6-102``."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _moments(b, hw, seed=0):
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(b, hw, hw, 8, generator=g)
    m[..., 4:] = m[..., 4:] * 3.0 - 4.0      # log-variances around exp(-2), some clipped at the [-30, 20] rails
    m[0, 0, 0, 4] = -50.0
    m[0, 0, 1, 5] = 35.0
    return m


@pytest.mark.parametrize("b,hw", [(1, 8), (2, 16), (3, 64)])
def test_noisy_latents_match_oracle(b, hw):
    from ddpo_b200 import ops
    from oracle import diffusion as OD, scheduler as OS, threefry
    mom = _moments(b, hw)
    sample_rng = threefry.split(threefry.PRNGKey(5), 3)[1]
    ac = OS.create_state(OS.SD_CONFIG).alphas_cumprod
    noisy_r, noise_r, ts_r, lat_r = OD.make_inputs(mom.numpy(), sample_rng, ac)
    noise_rng, timestep_rng = threefry.split(sample_rng)
    key = lambda k: (int(k[0]), int(k[1]))
    ts = ops.threefry_randint(key(timestep_rng), b, 0, 1000)
    assert ts == list(ts_r)
    dev = "cuda"
    keys = ops.key_tensor([key(sample_rng), key(noise_rng)], dev)
    noise = torch.empty(b, 4, hw, hw, device=dev)
    noisy = torch.empty_like(noise)
    lat = torch.empty_like(noise)
    ops.rwr_noisy_latents(mom.to(dev), keys[0], keys[1], torch.tensor(ts, dtype=torch.int32, device=dev),
                          torch.as_tensor(np.asarray(ac, np.float32)).to(dev), noise, noisy, latents_out=lat)
    torch.cuda.synchronize()
    # same threefry bits; erfinv's log1pf differs from NumPy's by an ulp (same bound as test_threefry_normal_matches_oracle)
    np.testing.assert_allclose(noise.cpu().numpy(), noise_r, rtol=0, atol=2e-6)
    np.testing.assert_allclose(lat.cpu().numpy(), lat_r, rtol=2e-6, atol=1e-6)     # expf vs np.exp: ~1 ulp
    np.testing.assert_allclose(noisy.cpu().numpy(), noisy_r, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("b,weighted,g", [(1, False, 5.0), (4, False, 5.0), (4, True, 5.0), (3, True, 0.0), (2, False, 1.0)])
def test_mse_loss_and_grad_match_oracle(b, weighted, g):
    from ddpo_b200 import ops
    from oracle import diffusion as OD
    n = 4 * 64 * 64
    gen = torch.Generator().manual_seed(b)
    eu = torch.randn(b, n, generator=gen).requires_grad_(True)
    ec = torch.randn(b, n, generator=gen).requires_grad_(True)
    noise = torch.randn(b, n, generator=gen)
    w = torch.rand(b, generator=gen) if weighted else None
    loss_r, per_r = OD.mse_loss(eu, ec, noise, g, True, w)
    loss_r.backward()
    dev = "cuda"
    ws = ops.rwr_workspace(b, dev)
    loss = torch.zeros(1, device=dev)
    per = torch.zeros(b, device=dev)
    du = torch.empty(b, n, device=dev)
    dc = torch.empty(b, n, device=dev)
    for _ in range(2):   # second call checks that the workspace counters reset themselves
        ops.rwr_mse_loss(eu.detach().to(dev), ec.detach().to(dev), noise.to(dev), g, loss, ws,
                         weights=None if w is None else w.to(dev), per_sample=per, d_eps_u=du, d_eps_c=dc)
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss.item(), loss_r.item(), rtol=1e-5)
    np.testing.assert_allclose(per.cpu().numpy(), per_r.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(du.cpu().numpy(), eu.grad.numpy(), rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(dc.cpu().numpy(), ec.grad.numpy(), rtol=1e-4, atol=1e-10)


def _setup(train_cfg, weighted, use_graph, b=2, lr=1e-3):
    from ddpo_b200 import unet_spec
    from ddpo_b200.diffusers_patch import DDIMScheduler
    from ddpo_b200.training import diffusion as D
    from ddpo_b200.unet import UNet
    D.USE_CUDA_GRAPH = use_graph
    D._GRAPHS.clear()
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(b, cfg.ctx_len, cfg.cross_attention_dim, generator=g)
    neg = torch.randn(1, cfg.ctx_len, cfg.cross_attention_dim, generator=g).expand(b, -1, -1).contiguous()
    net = UNet(cfg, flat, "cuda")
    sched = DDIMScheduler(1000, 0.00085, 0.012, "scaled_linear", device="cuda")
    st = sched.create_state()
    state = D.TrainState.create(apply_fn=net, tx=D.AdamWConfig(learning_rate=lr))
    mom = _moments(b, cfg.sample_size, seed=9)
    w = np.array([0.7, 0.3, 0.2, 0.1][:b], np.float32) if weighted else None
    batch = {"vae": mom, "prompt_embeds": emb, "uncond_embeds": neg}
    return cfg, flat, net, sched, st, state, batch, w, D


@pytest.mark.parametrize("train_cfg,weighted,use_graph", [(True, False, False), (True, True, True), (False, False, True)])
def test_rwr_train_step_matches_oracle(train_cfg, weighted, use_graph):
    from ddpo_b200 import unet_spec
    from oracle import diffusion as OD, optim as OO, scheduler as OS, threefry
    from oracle.unet import UNetOracle
    cfg, flat, net, sched, st, state, batch, w, D = _setup(train_cfg, weighted, use_graph)
    rng = (11, 22)
    state, loss, new_rng = D.train_step(state, None, batch, rng, st, (sched, None, train_cfg, 5.0), weights=w)
    torch.cuda.synchronize()
    # ---- oracle: same key lineage, same inputs
    _, sample_rng, new_r = OD.split3(np.array(rng, np.uint32))
    assert tuple(int(v) for v in new_r) == tuple(new_rng)
    ac = OS.create_state(OS.SD_CONFIG).alphas_cumprod
    noisy, noise, ts, _ = OD.make_inputs(batch["vae"].numpy(), sample_rng, ac)
    assert D.train_step.last["timesteps"] == list(ts)
    np.testing.assert_allclose(D.train_step.last["noise"].cpu().numpy(), noise, rtol=0, atol=2e-6)
    fp = flat.clone().requires_grad_(True)
    onet = UNetOracle(cfg, unet_spec.views(fp, cfg))
    loss_r, per_r = OD.train_loss(onet, (noisy, noise, ts), batch["prompt_embeds"], batch["uncond_embeds"], train_cfg,
                                  5.0, w)
    loss_r.backward()
    # the loss is O(1) (eps of a random-init net vs unit noise): bf16 U-Net error ~1e-2 relative on eps
    np.testing.assert_allclose(loss.item(), loss_r.item(), rtol=3e-2)
    # global gradient norm (pre-clip) and one clip+AdamW update from the oracle's gradient vs the CUDA parameters
    gref = fp.grad.numpy()
    ost = OO.AdamWState(flat.numel())
    p_ref, gn_ref = OO.clip_adamw_update(flat.numpy().copy(), gref, ost, lr=1e-3)
    np.testing.assert_allclose(state.last_grad_norm.item(), float(gn_ref), rtol=6e-2)
    p_new = net.params.cpu().numpy()
    # Adam's first step is ~ -lr * sign(g): compare the update direction where the oracle gradient is well above
    # the bf16 gradient error (~4e-2 of the norm, tests/test_gpu_training.py)
    big = np.abs(gref) > 0.1 * np.abs(gref).max()
    assert big.sum() > 10
    agree = np.sign(p_new - flat.numpy())[big] == np.sign(p_ref - flat.numpy())[big]
    assert agree.mean() > 0.95, agree.mean()
    # the gradient itself, as a norm bound: Adam's first moment after one step is (1 - b1) x the clipped gradient
    # (stored as bf16, 4e-3 relative) -> relative L2 and cosine against the oracle's clipped autograd gradient
    g_gpu = state.opt_state["mu"].float().cpu().numpy() / (1.0 - 0.9)
    g_clip = gref * min(1.0, 1.0 / float(gn_ref))
    rel = np.linalg.norm(g_gpu - g_clip) / np.linalg.norm(g_clip)
    cos = float(np.dot(g_gpu, g_clip) / (np.linalg.norm(g_gpu) * np.linalg.norm(g_clip)))
    assert rel < 8e-2 and cos > 0.997, (rel, cos)
    assert state.step == 1 and state.n_acc == 0


def test_rwr_graph_replay_equals_eager_and_rng_advances():
    """same rng + same batch -> identical loss with and without CUDA-graph capture; a new rng changes the draw"""
    losses = []
    for use_graph in (False, True, True):
        cfg, flat, net, sched, st, state, batch, w, D = _setup(True, False, use_graph, lr=0.0)
        state, loss, rng2 = D.train_step(state, None, batch, (1, 2), st, (sched, None, True, 5.0))
        losses.append(loss.item())
        state, loss_b, rng3 = D.train_step(state, None, batch, rng2, st, (sched, None, True, 5.0))
        assert rng3 != rng2 and loss_b.item() != loss.item()
    assert losses[0] == losses[1] == losses[2]
