"""Diagnostic (not a test): per-layer activation-gradient comparison CUDA vs oracle autograd."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_training import _sample, _batch
from ddpo_b200 import unet_spec
from ddpo_b200.training import policy_gradient as pg
from oracle import pipeline as OP, scheduler as OS
from oracle.unet import UNetOracle

pg.USE_CUDA_GRAPH = False
cfg, flat, emb, neg, net, sched, st, out = _sample()
state = pg.AccumulatingTrainState(apply_fn=net)
batch = _batch(out, emb, neg, 1, [1.5, -0.7])
net._dbg = {}
state, info = pg.train_step(state, batch, st, sched, True, 5.0, 1.0, 1e-4, False)
torch.cuda.synchronize()
G = list(pg._GRAPHS.values())[0]
d_eps_gpu = G.d_eps.cpu()
print("info", {k: v.item() for k, v in info.items()}, "logp", G.logp.cpu().numpy(), "dlogp", G.dlogp.cpu().numpy())

fp = flat.clone().requires_grad_(True)
taps = {}
def tap(n, t):
    if t.requires_grad:
        t.retain_grad()
    taps[n] = t
onet = UNetOracle(cfg, unet_spec.views(fp, cfg), tap=tap)
ost = OS.set_timesteps(OS.SD_CONFIG, OS.create_state(OS.SD_CONFIG), 3)
nb = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in batch.items()}
# run cond+uncond as one batch of 4 like the CUDA path to get per-layer grads in the same layout
lat = torch.as_tensor(nb["latents"]); ts = torch.as_tensor(np.asarray(nb["ts"])).long()
ctx = torch.cat([torch.as_tensor(nb["uncond_embeds"]), torch.as_tensor(nb["prompt_embeds"])])
eps_all = onet(torch.cat([lat, lat]), torch.cat([ts, ts]), ctx)
eps_all.retain_grad()
unc, cond = eps_all[:2], eps_all[2:]
eps = unc + 5.0 * (cond - unc)
a_t, a_prev, sigma = OS.coefficients(OS.SD_CONFIG, ost, np.asarray(nb["ts"]), 1.0)
tt = lambda v: torch.as_tensor(v, dtype=torch.float32).view(-1, 1, 1, 1)
a_t, a_prev, sigma = tt(a_t), tt(a_prev), tt(sigma)
x0 = (lat - torch.sqrt(1 - a_t) * eps) / torch.sqrt(a_t)
mean = torch.sqrt(a_prev) * x0 + torch.sqrt(1 - a_prev - sigma ** 2) * eps
lp = (-((torch.as_tensor(nb["next_latents"]) - mean) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - float(np.log(np.sqrt(2 * np.pi)))).mean(dim=(1, 2, 3))
adv = torch.tensor([1.5, -0.7])
ratio = torch.exp(lp - lp.detach())
loss = torch.maximum(-adv * ratio, -adv * torch.clamp(ratio, 1 - 1e-4, 1 + 1e-4)).mean()
loss.backward()
print("oracle lp", lp.detach().numpy(), "loss", loss.item())
def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
print("d_eps rel err", rel(d_eps_gpu, eps_all.grad), "norms", d_eps_gpu.norm().item(), eps_all.grad.norm().item())
for i in range(4):
    print("  sample", i, rel(d_eps_gpu[i], eps_all.grad[i]), d_eps_gpu[i].norm().item(), eps_all.grad[i].norm().item())
names = [n for n in taps if n in net._dbg]
for n in reversed(names):
    g = taps[n].grad
    if g is None:
        continue
    gg = net._dbg[n].cpu().view(g.shape)
    print(f"{n:60s} {rel(gg, g):.3e}  |ref|={g.norm().item():.3e}")
table, _ = unet_spec.param_offsets(cfg)
for name in ["conv_out/bias", "conv_out/kernel", "conv_norm_out/scale", "conv_norm_out/bias", "up_blocks_3/resnets_2/conv2/bias", "up_blocks_3/resnets_2/conv2/kernel"]:
    off, shape = table[name]; n = int(np.prod(shape))
    a, r = net.grads[off:off + n].cpu(), fp.grad[off:off + n]
    print(name, rel(a, r), a.norm().item(), r.norm().item(), (a[:4] / r[:4]).numpy())
