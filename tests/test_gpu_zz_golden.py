"""CUDA path against the committed golden fixtures (tests/golden/make_golden_r1b.py): RWR kernels and the VAE decoder
against the oracle's frozen outputs, the text encoder against vectors produced by transformers' own CLIPTextModel."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.gpu
def test_gpu_rwr_golden():
    from ddpo_b200 import ops
    from oracle import scheduler as OS, threefry
    z = np.load(os.path.join(G, "rwr.npz"))
    dev = "cuda"
    key = lambda k: (int(k[0]), int(k[1]))
    _, srng, new = ops.threefry_split(key(z["train_rng"]), 3)
    assert srng == key(z["sample_rng"]) and new == key(z["new_rng"])
    nrng, trng = ops.threefry_split(srng, 2)
    ts = ops.threefry_randint(trng, 3, 0, 1000)
    assert ts == list(z["timesteps"])
    keys = ops.key_tensor([srng, nrng], dev)
    noise = torch.empty(3, 4, 8, 8, device=dev)
    noisy, lat = torch.empty_like(noise), torch.empty_like(noise)
    ac = torch.as_tensor(np.asarray(OS.create_state(OS.SD_CONFIG).alphas_cumprod, np.float32)).to(dev)
    ops.rwr_noisy_latents(torch.from_numpy(z["moments"]).to(dev), keys[0], keys[1],
                          torch.tensor(ts, dtype=torch.int32, device=dev), ac, noise, noisy, latents_out=lat)
    loss = torch.zeros(1, device=dev)
    per = torch.zeros(3, device=dev)
    ops.rwr_mse_loss(torch.from_numpy(z["eps_u"]).to(dev), torch.from_numpy(z["eps_c"]).to(dev),
                     torch.from_numpy(z["noise"].reshape(3, -1)).to(dev), 5.0, loss, ops.rwr_workspace(3, dev),
                     weights=torch.from_numpy(z["weights"]).to(dev), per_sample=per,
                     d_eps_u=torch.empty(3, 256, device=dev), d_eps_c=torch.empty(3, 256, device=dev))
    torch.cuda.synchronize()
    np.testing.assert_allclose(noise.cpu().numpy(), z["noise"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(noisy.cpu().numpy(), z["noisy"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    np.testing.assert_allclose(per.cpu().numpy(), z["per_sample"], rtol=1e-5)


@pytest.mark.gpu
def test_gpu_vae_golden():
    from ddpo_b200 import vae as V
    z = np.load(os.path.join(G, "vae_micro.npz"))
    dec = V.VAEDecoder(V.VAE_MICRO, V.init_flat_params(V.VAE_MICRO, 0), "cuda")
    raw = dec.decode(torch.from_numpy(z["latents"]).cuda(), want_raw=True, want_images=False)
    torch.cuda.synchronize()
    rel = np.linalg.norm(raw.cpu().numpy() - z["raw"]) / np.linalg.norm(z["raw"])
    assert rel < 3e-2, rel


@pytest.mark.gpu
@pytest.mark.parametrize("name,cfg_name", [("gelu", "TEXT_TINY"), ("quick_gelu", "TEXT_TINY_QUICK")])
def test_gpu_text_encoder_golden_from_transformers(name, cfg_name):
    from ddpo_b200 import text_encoder as T
    z = np.load(os.path.join(G, f"text_tiny_{name}.npz"))
    cfg = getattr(T, cfg_name)
    enc = T.CLIPTextEncoder(cfg, T.init_flat_params(cfg, 0), "cuda")
    got = enc(z["input_ids"])[0]
    torch.cuda.synchronize()
    ref = z["last_hidden_state"]
    rel = np.linalg.norm(got.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert rel < 2e-2, rel
