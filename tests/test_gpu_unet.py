"""U-Net forward on the CUDA path vs the CPU oracle (small configs the oracle finishes in seconds)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(cfg_name, batch, seed=0):
    from ddpo_b200 import unet_spec
    from ddpo_b200.unet import UNet
    from oracle.unet import UNetOracle
    cfg = getattr(unet_spec, cfg_name)
    flat = unet_spec.init_flat_params(cfg, seed)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    s = cfg.sample_size
    lat = torch.randn(batch, 4, s, s, generator=g)
    ctx = torch.randn(batch, cfg.ctx_len, cfg.cross_attention_dim, generator=g)
    ts = torch.tensor([981, 441, 21, 1][:batch], dtype=torch.int32)
    ref_taps = {}
    ref = UNetOracle(cfg, unet_spec.views(flat, cfg), torch.float32, tap=lambda n, t: ref_taps.__setitem__(n, t))(
        lat, ts, ctx)
    net = UNet(cfg, flat, "cuda")
    net.prepare_context(ctx.cuda())
    taps = {}
    out = net.forward(lat.cuda(), ts.cuda(), taps=taps)
    torch.cuda.synchronize()
    report = []
    for name, t in taps.items():
        if name in ref_taps:
            r = ref_taps[name]
            e = ((t.cpu().float() - r).norm() / (r.norm() + 1e-12)).item()
            report.append((name, e))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/unet_taps_{cfg_name}.txt", "w") as f:
        for name, e in report:
            f.write(f"{name}\t{e:.3e}\n")
    err = ((out.cpu() - ref).norm() / ref.norm()).item()
    return err, report, out, ref


@pytest.mark.parametrize("cfg_name,batch", [("TINY", 2), ("TINY", 3), ("SMALL", 2)])
def test_unet_forward_matches_oracle(cfg_name, batch):
    err, report, out, ref = _run(cfg_name, batch)
    worst = max(report, key=lambda x: x[1]) if report else None
    print(f"{cfg_name} b={batch}: eps rel err {err:.3e}; worst tap {worst}")
    assert err < 3e-2, f"eps relative error {err} (worst layer {worst})"


def test_unet_batch_invariance():
    """Rows of a sample must be bit-identical whether it is run alone or inside a larger batch
    (first-pass PPO ratio == 1; reference config/base.py:88 ppo_clip_range 1e-4)."""
    from ddpo_b200 import unet_spec
    from ddpo_b200.unet import UNet
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator(device="cpu").manual_seed(5)
    lat = torch.randn(4, 4, 16, 16, generator=g).cuda()
    ctx = torch.randn(4, cfg.ctx_len, cfg.cross_attention_dim, generator=g).cuda()
    ts = torch.tensor([981, 981, 441, 441], dtype=torch.int32).cuda()
    net = UNet(cfg, flat, "cuda")
    net.prepare_context(ctx)
    big = net.forward(lat, ts).clone()
    net.prepare_context(ctx[2:4].contiguous())
    small = net.forward(lat[2:4].contiguous(), ts[2:4].contiguous()).clone()
    assert torch.equal(big[2:4], small)
