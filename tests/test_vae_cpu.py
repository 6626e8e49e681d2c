"""CPU tests of the VAE-decoder spec and its oracle restatement (self-consistency pins; no diffusers offline)."""
import numpy as np
import torch

from ddpo_b200 import vae as V
from oracle import vae as OV


def test_sd_decoder_parameter_count():
    """AutoencoderKL of Stable Diffusion: 83 653 863 parameters in total, 34 163 592 in the encoder (+ quant_conv 72);
    decoder + post_quant_conv = 49 490 199 -- the manifest must reproduce the published decoder size."""
    n = V.num_params(V.SD_VAE)
    assert n == 49_490_199, n


def test_manifest_names_and_shapes():
    names = dict(V.param_manifest(V.SD_VAE))
    assert names["decoder/conv_in/kernel"] == (3, 3, 4, 512)
    assert names["decoder/up_blocks_2/resnets_0/conv_shortcut/kernel"] == (1, 1, 512, 256)
    assert names["decoder/up_blocks_3/resnets_0/conv_shortcut/kernel"] == (1, 1, 256, 128)
    assert "decoder/up_blocks_3/upsamplers_0/conv/kernel" not in names
    assert names["decoder/mid_block/attentions_0/proj_attn/kernel"] == (512, 512)
    assert names["decoder/conv_out/kernel"] == (3, 3, 128, 3)
    table, total = V.param_offsets(V.SD_VAE)
    assert all(off % 64 == 0 for off, _ in table.values()) and total >= V.num_params(V.SD_VAE)


def test_oracle_decode_shapes_range_and_linearity_of_head():
    cfg = V.VAE_MICRO
    flat = V.init_flat_params(cfg, 0)
    p = V.views(flat, cfg)
    lat = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0)) * 0.18215
    taps = {}
    img, raw = OV.decode(p, cfg, lat, taps=taps)
    assert img.shape == (2, 64, 64, 3) and raw.shape == (2, 3, 64, 64)
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    np.testing.assert_allclose(img.numpy(), np.clip(raw.permute(0, 2, 3, 1).numpy() / 2 + 0.5, 0, 1), atol=1e-7)
    assert taps["mid"].shape == (2, 8, 8, 128) and taps["up2"].shape == (2, 64, 64, 64)
    # batch independence: decoding a sample alone gives the same image
    img1, _ = OV.decode(p, cfg, lat[1:])
    np.testing.assert_allclose(img1.numpy(), img[1:].numpy(), atol=2e-4)
    # fp64 truth close to fp32
    img64, _ = OV.decode(p, cfg, lat, dtype=torch.float64)
    assert float((img64 - img).abs().max()) < 1e-4


def test_attention_uniform_when_keys_equal():
    """with zero query/key weights every score is equal -> the block returns proj(mean_pixels(value)) + x"""
    cfg = V.VAE_MICRO
    flat = V.init_flat_params(cfg, 1)
    p = {k: v.clone() for k, v in V.views(flat, cfg).items()}
    name = "decoder/mid_block/attentions_0"
    p[name + "/query/kernel"].zero_()
    p[name + "/query/bias"].zero_()
    x = torch.randn(1, 4, 4, 128, generator=torch.Generator().manual_seed(2))
    y = OV._attention(x, p, name)
    from oracle.unet import group_norm
    g = group_norm(x, p[name + "/group_norm/scale"], p[name + "/group_norm/bias"], eps=1e-6).reshape(1, 16, 128)
    v = g @ p[name + "/value/kernel"] + p[name + "/value/bias"]
    want = (v.mean(1, keepdim=True) @ p[name + "/proj_attn/kernel"] + p[name + "/proj_attn/bias"]).reshape(1, 1, 1, 128) + x
    np.testing.assert_allclose(y.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def test_vae_host_assembly_dry_run_against_oracle(monkeypatch):
    """ddpo_b200/vae.py's kernel sequencing, dry-run on a torch-CPU emulation of the ops it calls (tests/
    _cpu_ops_emulator.py; same bf16-operand / fp32-accumulate conventions): buffer shapes, operand order, bias and
    residual plumbing, the V^T trick of the attention block, chunked decode.  The CUDA kernels themselves are checked
    on the GPU box (tests/test_gpu_z_vae.py)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import _cpu_ops_emulator as E
    monkeypatch.setattr(V, "ops", E)
    monkeypatch.setattr(V, "Arena", E.CpuArena)
    for cfg, b, db in ((V.VAE_MICRO, 3, 2), (V.VAEConfig(block_out_channels=(64, 64, 128, 128), sample_size=16), 1, 1)):
        flat = V.init_flat_params(cfg, 0)
        dec = V.VAEDecoder(cfg, flat, device="cpu", decode_batch=db)
        lat = torch.randn(b, 4, cfg.sample_size, cfg.sample_size, generator=torch.Generator().manual_seed(4)) * 0.18215
        img, raw = dec.decode(lat, want_raw=True, want_images=True)
        img_r, raw_r = OV.decode(V.views(flat, cfg), cfg, lat)
        rel = ((raw - raw_r).norm() / raw_r.norm()).item()
        assert rel < 3e-2, rel
        assert (img - img_r).abs().max().item() < 0.1


def test_vae_encoder_host_assembly_dry_run_against_oracle(monkeypatch):
    """ddpo_b200/vae.py::VAEEncoder's kernel sequencing on the CPU ops emulator against oracle/vae.py::encode: image
    normalisation, conv_in, the one-sided stride-2 down-sampling, mid block, fp32 head (conv_out, quant_conv, logvar clip)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import _cpu_ops_emulator as E
    monkeypatch.setattr(V, "ops", E)
    monkeypatch.setattr(V, "Arena", E.CpuArena)
    cfg = V.VAE_MICRO
    flat = V.init_flat_params(cfg, 1, part="encoder")
    enc = V.VAEEncoder(cfg, flat, device="cpu", decode_batch=2)
    img = torch.rand(3, 64, 64, 3, generator=torch.Generator().manual_seed(8))
    mom = enc.encode(img)
    ref = OV.encode(V.views(flat, cfg, part="encoder"), cfg, img)
    assert mom.shape == ref.shape == (3, 8, 8, 8)
    assert ((mom - ref).norm() / ref.norm()).item() < 3e-2
