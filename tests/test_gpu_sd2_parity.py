"""Parity at the BASELINE configuration: the SD2-base U-Net (866 M parameters, C = 320..1280, 20 heads, GEGLU N = 10240,
two-source convolutions up to K = 23040) on the CUDA path against the fp32 CPU oracle -- one denoising step (reference
``pipeline_flax_stable_diffusion.py:204-241``) and one PPO train step (``training/policy_gradient.py:86-138``) of ONE
sample -- plus the kernel shapes only the full model reaches, against plain PyTorch fp32 on the same bf16 operands.

Measured numbers are appended to ``gpurun_out/sd2_parity.txt`` (DESIGN.md section 4 quotes them).
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
G, ETA, T = 5.0, 1.0, 50


def _log(line):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/sd2_parity.txt", "a") as f:
        f.write(line + "\n")
    print(line)


def _rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module")
def sd2():
    """SD2-base on both sides from the same seeded CPU weights; the oracle's first denoising step is computed once."""
    from ddpo_b200 import unet_spec
    from ddpo_b200.diffusers_patch import DDIMScheduler
    from ddpo_b200.unet import UNet
    from oracle import ppo as OPPO, scheduler as OS, threefry
    from oracle.unet import UNetOracle
    import psutil
    # physical cores: one thread per hyper-thread makes oneDNN's backward path collapse on the 64-core / 128-thread box
    torch.set_num_threads(max(1, min(psutil.cpu_count(logical=False) or 8, len(os.sched_getaffinity(0)))))
    cfg = unet_spec.SD2_BASE
    flat = unet_spec.init_flat_params(cfg, 0)
    onet = UNetOracle(cfg, unet_spec.views(flat, cfg))
    g = torch.Generator().manual_seed(1)
    ctx = torch.randn(2, 77, 1024, generator=g)                                  # [uncond ; cond]
    x = threefry.normal(np.array((0, 7), np.uint32), (1, 4, 64, 64)).astype(np.float32)
    key = (0, 11)
    ost = OS.set_timesteps(OS.SD_CONFIG, OS.create_state(OS.SD_CONFIG), T)
    t0 = int(ost.timesteps[0])
    xt = torch.from_numpy(x)
    with torch.no_grad():
        e = onet(torch.cat([xt, xt]), torch.full((2,), t0), ctx).numpy()
    eps = OPPO.cfg_combine(e[:1], e[1:], G)
    prev, _, lp = OS.step(OS.SD_CONFIG, ost, eps, t0, x, key=np.array(key, np.uint32), eta=ETA)
    net = UNet(cfg, flat, DEV)
    sched = DDIMScheduler(1000, 0.00085, 0.012, "scaled_linear", None, False, 1, "epsilon", device=DEV)
    st = sched.set_timesteps(sched.create_state(), T)
    return dict(cfg=cfg, flat=flat, onet=onet, ctx=ctx, x=x, key=key, ost=ost, t0=t0, e=e, prev=prev, lp=np.asarray(lp),
                net=net, sched=sched, st=st)


def test_sd2_base_denoising_step_matches_oracle(sd2):
    from ddpo_b200 import ops
    s = sd2
    net, st = s["net"], s["st"]
    xg = torch.from_numpy(s["x"]).to(DEV).view(1, -1)
    tdev = torch.tensor([s["t0"]], dtype=torch.int32, device=DEV)
    net.prepare_context(s["ctx"].to(DEV))
    taps = {}
    eps_g = net.forward(torch.cat([xg, xg]).view(2, 4, 64, 64), tdev, taps=taps).view(2, -1)
    prev_g, lp_g, lp_s = torch.empty_like(xg), torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    ws = ops.ddim_workspace(1, DEV)
    ac, fa = st.common.alphas_cumprod, st.final_alpha_cumprod
    ops.ddim_step_sample(eps_g[:1], eps_g[1:], xg, ac, tdev, fa, 20, G, ETA, ops.key_tensor([s["key"]], DEV), prev_g, lp_g, ws)
    ops.ddim_logprob_fwd(eps_g[:1], eps_g[1:], xg, torch.from_numpy(s["prev"]).to(DEV).view(1, -1), ac, tdev, fa, 20, G, ETA,
                         lp_s, ws)
    torch.cuda.synchronize()
    eps_rel = _rel(eps_g.cpu().numpy(), s["e"])
    lat_rel = _rel(prev_g.cpu().numpy(), s["prev"])
    lp_o = float(s["lp"].ravel()[0])
    lp_rel = abs(lp_g.item() / lp_o - 1)
    lps_rel = abs(lp_s.item() / lp_o - 1)
    ratio = math.exp(lp_s.item() - lp_o)
    _log(f"SD2-base denoise step t={s['t0']}: eps rel-L2 {eps_rel:.3e}  first-step latents rel-L2 {lat_rel:.3e}  "
         f"log_prob (sample mode) rel {lp_rel:.3e}  log_prob of the oracle's x_prev under GPU eps rel {lps_rel:.3e}  "
         f"importance ratio {ratio:.6f}  (log_prob oracle {lp_o:.6f})")
    assert np.isfinite(eps_g.cpu().numpy()).all()
    assert eps_rel < 2.5e-2, f"eps relative L2 error {eps_rel}"
    assert lat_rel < 2.5e-2, f"first-step latents relative L2 error {lat_rel}"
    assert lp_rel < 1e-3, f"per-step log_prob relative error {lp_rel} (north-star tolerance 1e-3)"
    # the score-mode number folds the bf16 U-Net error through (x_prev - mean)/sigma^2: reported, loosely bounded
    assert lps_rel < 2e-2 and abs(ratio - 1) < 2e-2, (lps_rel, ratio)


def test_sd2_base_batch_invariance_of_eps(sd2):
    """The sample's eps is bit-identical inside a larger batch (what makes ratio == 1 at the reference's clip range 1e-4)."""
    s = sd2
    net = s["net"]
    xg = torch.from_numpy(s["x"]).to(DEV).view(1, 4, 64, 64)
    tdev = torch.tensor([s["t0"]], dtype=torch.int32, device=DEV)
    net.prepare_context(s["ctx"].to(DEV))
    e2 = net.forward(torch.cat([xg, xg]), tdev).clone()
    g = torch.Generator().manual_seed(9)
    other = torch.randn(2, 4, 64, 64, generator=g).to(DEV)
    octx = torch.randn(2, 77, 1024, generator=g).to(DEV)
    # batch 6 = [u(x), u(o0), u(o1), c(x), c(o0), c(o1)]
    ctx6 = torch.cat([s["ctx"][:1].to(DEV), octx, s["ctx"][1:].to(DEV), octx])
    net.prepare_context(ctx6)
    e6 = net.forward(torch.cat([xg, other, xg, other]), tdev)
    torch.cuda.synchronize()
    assert torch.equal(e6[0], e2[0]) and torch.equal(e6[3], e2[1])


@pytest.mark.timeout(420)
def test_sd2_base_ppo_train_step_matches_oracle(sd2):
    """loss, log-prob, ratio and the parameter gradient of one reference-sized train step (train_cfg, batch 1)."""
    from ddpo_b200 import unet_spec
    from ddpo_b200.training import policy_gradient as pg
    from oracle import pipeline as OP, scheduler as OS
    from oracle.unet import UNetOracle
    s = sd2
    cfg = s["cfg"]
    clip = 1e9   # both sides on the unclipped branch -A * ratio whatever their own ratio is (|ratio - 1| ~ 1e-3)
    adv = 1.0
    batch_np = {"latents": s["x"], "next_latents": s["prev"], "ts": np.array([s["t0"]], np.int32),
                "log_probs": s["lp"].reshape(1).astype(np.float32), "advantages": np.array([adv], np.float32),
                "prompt_embeds": s["ctx"][1:].numpy(), "uncond_embeds": s["ctx"][:1].numpy()}
    params = {k: v.clone().requires_grad_(True) for k, v in unet_spec.views(s["flat"], cfg).items()}
    onet = UNetOracle(cfg, params)
    loss, info, lp = OP.train_loss(onet, OS.SD_CONFIG, s["ost"], batch_np, True, G, ETA, clip)
    loss.backward()
    ref = {k: p.grad for k, p in params.items()}
    # CUDA
    pg.USE_CUDA_GRAPH = False
    pg._GRAPHS.clear()
    net = s["net"]
    state = pg.AccumulatingTrainState(apply_fn=net)
    net.grads.zero_()
    batch = {k: torch.as_tensor(v).to(DEV) for k, v in batch_np.items()}
    state, ginfo = pg.train_step(state, batch, s["st"], s["sched"], True, G, ETA, clip, False)
    torch.cuda.synchronize()
    gg = net.grads.cpu()
    table, _ = unet_spec.param_offsets(cfg)
    tot_num = tot_den = dot = na = nb = 0.0
    blocks = {}
    worst, worst_name = 0.0, None
    ref_total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in ref.values()))
    for name, (off, shape) in table.items():
        n = int(np.prod(shape))
        a, r = gg[off:off + n].double(), ref[name].reshape(-1).double()
        d2, r2 = float(((a - r) ** 2).sum()), float((r ** 2).sum())
        tot_num += d2
        tot_den += r2
        dot += float((a * r).sum())
        na += float((a ** 2).sum())
        nb += r2
        blk = name.split("/")[0]
        b = blocks.setdefault(blk, [0.0, 0.0])
        b[0] += d2
        b[1] += r2
        if math.sqrt(r2) > 1e-3 * ref_total / math.sqrt(len(table)):
            e = math.sqrt(d2 / r2)
            if e > worst:
                worst, worst_name = e, name
    grad_rel = math.sqrt(tot_num / tot_den)
    cos = dot / math.sqrt(na * nb)
    norm_ratio = math.sqrt(na / nb)
    _log(f"SD2-base PPO train step: loss gpu {ginfo['loss'].item():.6f} oracle {float(loss):.6f}  "
         f"approx_kl gpu {ginfo['approx_kl'].item():.3e}  grad rel-L2 {grad_rel:.3e}  cosine {cos:.6f}  "
         f"|g_gpu|/|g_oracle| {norm_ratio:.4f}  worst tensor {worst_name} {worst:.3e}")
    for blk, (d2, r2) in blocks.items():
        _log(f"    block {blk:24s} grad rel-L2 {math.sqrt(d2 / r2):.3e}   |g| {math.sqrt(r2):.3e}")
    assert abs(ginfo["loss"].item() - float(loss)) < 2e-2 * max(1.0, abs(float(loss)))
    assert cos > 0.995 and abs(norm_ratio - 1) < 5e-2, (cos, norm_ratio)
    assert grad_rel < 0.1, grad_rel
    for blk, (d2, r2) in blocks.items():
        assert math.sqrt(d2 / r2) < 0.2, (blk, math.sqrt(d2 / r2))
    # leave the shared network as it was found
    net.grads.zero_()
    pg.USE_CUDA_GRAPH = True


# ------------------------------------------------- kernel shapes only the full model reaches -----------------
def bf(x):
    return x.to(torch.bfloat16)


def _prep_w(w_kn):
    from ddpo_b200 import ops
    k, n = w_kn.shape
    dst = torch.empty(n, k, dtype=torch.bfloat16, device=DEV)
    ops.prep_weight(w_kn.contiguous(), dst, k, n)
    return dst


def _conv_ref(x_nhwc, w_hwio, bias, stride=1):
    y = torch.nn.functional.conv2d(x_nhwc.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), bias, stride=stride,
                                   padding=w_hwio.shape[0] // 2)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("b,h,c0,c1,n", [(2, 8, 1280, 1280, 1280),     # up_blocks_0: K = 23040, M = 64 B on igemm_kernel
                                         (2, 16, 1280, 640, 1280),     # up_blocks_1/resnets_2: K = 17280
                                         (1, 32, 640, 320, 640),       # up_blocks_2/resnets_2: K = 8640
                                         (1, 64, 320, 320, 320)])      # up_blocks_3: K = 5760, M = 4096 B
def test_two_source_conv_at_sd2_shapes(b, h, c0, c1, n):
    from ddpo_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(20)
    x0 = bf(torch.randn(b, h, h, c0, generator=g)).to(DEV)
    x1 = bf(torch.randn(b, h, h, c1, generator=g)).to(DEV)
    cin = c0 + c1
    w = (torch.randn(3, 3, cin, n, generator=g) / math.sqrt(9 * cin)).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV)
    res = torch.randn(b * h * h, n, generator=g).to(DEV)
    wt = _prep_w(w.reshape(9 * cin, n))
    out = torch.zeros(b * h * h, n, device=DEV)
    ops.igemm(a0=x0, a1=x1, wt=wt, n=n, c0=c0, c1=c1, conv=(b, h, h), taps=9, bias=bias, residual=res, out_f32=out)
    torch.cuda.synchronize()
    ref = _conv_ref(torch.cat([x0.float(), x1.float()], -1), bf(w).float(), bias).reshape(b * h * h, n) + res
    err = (out - ref).abs().max().item()
    assert err < 3e-3 * max(1.0, ref.abs().max().item()), f"max err {err}"
    # weight gradient of the same layer through the split planner
    dy = bf(torch.randn(b * h * h, n, generator=g) / 8).to(DEV)
    dw = torch.zeros(9 * cin, n, device=DEV)
    ops.wgrad(dy=dy, n=n, x0=x0, x1=x1, c0=c0, c1=c1, conv=(b, h, h), taps=9, dw=dw)
    torch.cuda.synchronize()
    xin = torch.cat([x0.float(), x1.float()], -1).permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.nn.grad.conv2d_weight(xin, (n, cin, 3, 3), dy.float().view(b, h, h, n).permute(0, 3, 1, 2), padding=1)
    wref = wref.permute(2, 3, 1, 0).reshape(9 * cin, n)
    rel = ((dw - wref).norm() / wref.norm()).item()
    assert rel < 2e-3, rel


@pytest.mark.parametrize("m,c", [(256, 1280), (128, 1280), (4096, 320)])
def test_geglu_at_sd2_shapes(m, c):
    """GEGLU up-projection N = 8C = 10240 (C = 1280) / 2560 and the FF down-projection K = 4C."""
    from ddpo_b200 import ops
    g = torch.Generator().manual_seed(21)
    a = bf(torch.randn(m, c, generator=g)).to(DEV)
    w = (torch.randn(c, 8 * c, generator=g) / math.sqrt(c)).to(DEV)
    bias = torch.randn(8 * c, generator=g).to(DEV)
    wt = torch.empty(8 * c, c, dtype=torch.bfloat16, device=DEV)
    ops.prep_weight(w, wt, c, 8 * c, geglu_bn=256)
    bp = torch.empty_like(bias)
    ops.permute_geglu_bias(bias, bp, 8 * c, 256)
    out = torch.zeros(m, 4 * c, dtype=torch.bfloat16, device=DEV)
    ops.igemm(a0=a, wt=wt, n=8 * c, c0=c, m=m, bias=bp, out_bf16=out, geglu=True, bn=256)
    f = a.float() @ bf(w).float() + bias
    lin, gate = f.chunk(2, dim=-1)
    ref = lin * torch.nn.functional.gelu(gate, approximate="tanh")
    assert ((out.float() - ref).norm() / ref.norm()).item() < 6e-3
    w2 = (torch.randn(4 * c, c, generator=g) / math.sqrt(4 * c)).to(DEV)
    res = torch.randn(m, c, generator=g).to(DEV)
    o2 = torch.zeros(m, c, device=DEV)
    ops.igemm(a0=out, wt=_prep_w(w2), n=c, c0=4 * c, m=m, residual=res, out_f32=o2)
    ref2 = out.float() @ bf(w2).float() + res
    assert (o2 - ref2).abs().max().item() < 3e-3 * max(1.0, ref2.abs().max().item())


@pytest.mark.parametrize("b,heads,nq,nk", [(2, 20, 256, 256), (2, 20, 64, 64), (2, 20, 256, 77), (2, 20, 64, 77),
                                           (1, 10, 1024, 1024), (1, 10, 1024, 77), (1, 5, 4096, 77)])
def test_attention_at_sd2_head_counts(b, heads, nq, nk):
    from ddpo_b200 import ops
    g = torch.Generator().manual_seed(22)
    c = heads * 64
    q = bf(torch.randn(b, nq, c, generator=g)).to(DEV)
    kv = bf(torch.randn(b, nk, 2 * c, generator=g)).to(DEV)
    out = torch.zeros(b, nq, c, dtype=torch.bfloat16, device=DEV)
    lse = torch.zeros(b, heads, nq, device=DEV)
    ops.attention_fwd(q, kv, kv[:, :, c:], out, b, heads, nq, nk, c, 2 * c, 2 * c, c, lse=lse)
    qh = q.float().view(b, nq, heads, 64).permute(0, 2, 1, 3)
    kh = kv[:, :, :c].float().reshape(b, nk, heads, 64).permute(0, 2, 1, 3)
    vh = kv[:, :, c:].float().reshape(b, nk, heads, 64).permute(0, 2, 1, 3)
    sc = (qh @ kh.transpose(-1, -2)) * 0.125
    p = torch.softmax(sc, dim=-1)
    ref = (p @ vh).permute(0, 2, 1, 3).reshape(b, nq, c)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 1e-2
    assert (lse - torch.logsumexp(sc, dim=-1)).abs().max().item() < 1e-3
    # backward at the same shape
    do = bf(torch.randn(b, nq, c, generator=g)).to(DEV)
    dq = torch.zeros(b, nq, c, dtype=torch.bfloat16, device=DEV)
    dkv = torch.zeros(b, nk, 2 * c, dtype=torch.bfloat16, device=DEV)
    delta = torch.zeros(b, heads, nq, device=DEV)
    ops.attention_bwd(q, kv, kv[:, :, c:], out, do, lse, delta, dq, dkv, dkv[:, :, c:], b, heads, nq, nk, c, 2 * c, 2 * c, c,
                      c, c, 2 * c, 2 * c)
    doh = do.float().view(b, nq, heads, 64).permute(0, 2, 1, 3)
    dv = p.transpose(-1, -2) @ doh
    dp = doh @ vh.transpose(-1, -2)
    ds = p * (dp - (dp * p).sum(-1, keepdim=True)) * 0.125
    dqr = (ds @ kh).permute(0, 2, 1, 3).reshape(b, nq, c)
    dkr = (ds.transpose(-1, -2) @ qh).permute(0, 2, 1, 3).reshape(b, nk, c)
    dvr = dv.permute(0, 2, 1, 3).reshape(b, nk, c)
    for got, want, nm in ((dq, dqr, "dq"), (dkv[:, :, :c], dkr, "dk"), (dkv[:, :, c:], dvr, "dv")):
        e = ((got.float() - want).norm() / want.norm()).item()
        assert e < 2e-2, (nm, e)
