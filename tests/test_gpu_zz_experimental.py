"""Default-OFF optimisations written after this round's GPU budget was spent (they have never run on a GPU): each is
selected by an environment variable and checked here for BIT-IDENTITY against the default path.  The whole module is
skipped unless DDPO_EXPERIMENTAL=1, so the round-end suite does not depend on unmeasured code; round 2 starts by running
    DDPO_EXPERIMENTAL=1 python -m pytest tests/test_gpu_zz_experimental.py -m gpu
and flips the defaults that pass and pay."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DDPO_EXPERIMENTAL") != "1", reason="experimental paths: set DDPO_EXPERIMENTAL=1")]
DEV = "cuda"


@pytest.mark.parametrize("batch", [1, 2, 16, 40])
def test_dense_small_grouped_is_bit_identical(batch):
    from ddpo_b200 import ops
    g = torch.Generator().manual_seed(0)
    k, ns = 256, [64, 320, 96, 640]
    x = torch.randn(batch, k, generator=g).to(DEV)
    params, entries, off = [], [], 0
    for n in ns:
        w, b = torch.randn(k, n, generator=g) / 16, torch.randn(n, generator=g)
        entries.append((off, off + k * n, None, n))
        params += [w.reshape(-1), b]
        off += k * n + n
    flat = torch.cat(params).to(DEV)
    y_off, fixed = 0, []
    for (w_off, b_off, _, n) in entries:
        fixed.append((w_off, b_off, y_off, n))
        y_off += batch * n
    table, ctas = ops.dense_small_group_table(fixed, DEV)
    y_all = torch.zeros(y_off, device=DEV)
    ops.dense_small_grouped(x, flat, y_all, table, len(ns), ctas, batch, k)
    torch.cuda.synchronize()
    for (w_off, b_off, yo, n) in fixed:
        ref = torch.empty(batch, n, device=DEV)
        ops.dense_small(x, flat[w_off:w_off + k * n].view(k, n), flat[b_off:b_off + n], ref, batch, k, n)
        torch.cuda.synchronize()
        assert torch.equal(y_all[yo:yo + batch * n].view(batch, n), ref)


@pytest.mark.parametrize("cfg_name,batch", [("TINY", 2), ("SMALL", 4)])
def test_unet_with_grouped_temb_is_bit_identical(cfg_name, batch, monkeypatch):
    from ddpo_b200 import unet_spec
    from ddpo_b200.unet import UNet
    cfg = getattr(unet_spec, cfg_name)
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(batch, 4, cfg.sample_size, cfg.sample_size, generator=g).to(DEV)
    ctx = torch.randn(batch, cfg.ctx_len, cfg.cross_attention_dim, generator=g).to(DEV)
    ts = torch.tensor([981, 441, 21, 1][:batch], dtype=torch.int32, device=DEV)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("DDPO_GROUPED_TEMB", flag)
        net = UNet(cfg, flat, DEV)
        assert net.grouped_temb == (flag == "1")
        net.prepare_context(ctx)
        outs.append(net.forward(lat, ts).clone())
        net.enable_training()      # and through the taped forward + backward
        tape = []
        eps = net.forward(lat, ts, tape=tape)
        net.backward(tape, torch.ones_like(eps) / eps.numel())
        outs.append(net.grads.clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])
