"""Golden fixtures of the components added after the first batch (tests/golden/make_golden_r1b.py): the oracle against
its frozen outputs (RWR, VAE) and against vectors produced by transformers' own CLIPTextModel (text encoder); the GPU
variants run the CUDA path against the same files."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), "golden")


def test_rwr_golden():
    from oracle import diffusion as OD, scheduler as OS, threefry
    z = np.load(os.path.join(G, "rwr.npz"))
    _, srng, new = OD.split3(z["train_rng"])
    assert np.array_equal(srng, z["sample_rng"]) and np.array_equal(new, z["new_rng"])
    noisy, noise, ts, lat = OD.make_inputs(z["moments"], srng, OS.create_state(OS.SD_CONFIG).alphas_cumprod)
    assert np.array_equal(ts, z["timesteps"])
    np.testing.assert_allclose(noise, z["noise"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(lat, z["latents"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(noisy, z["noisy"], rtol=1e-6, atol=1e-6)
    loss, per = OD.mse_loss(torch.from_numpy(z["eps_u"]), torch.from_numpy(z["eps_c"]),
                            torch.from_numpy(z["noise"].reshape(3, -1)), 5.0, True, z["weights"])
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-6)
    np.testing.assert_allclose(per.numpy(), z["per_sample"], rtol=1e-6)
    assert np.array_equal(threefry.randint(threefry.PRNGKey(5), (17,), 0, 1000), z["randint_17"])


def test_vae_golden():
    from ddpo_b200 import vae as V
    from oracle import vae as OV
    z = np.load(os.path.join(G, "vae_micro.npz"))
    flat = V.init_flat_params(V.VAE_MICRO, 0)
    assert abs(flat.double().sum().item() - float(z["param_sum"])) < 1e-6
    img, raw = OV.decode(V.views(flat, V.VAE_MICRO), V.VAE_MICRO, z["latents"])
    np.testing.assert_allclose(raw.numpy(), z["raw"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(img.mean(dim=(1, 2)).numpy(), z["image_mean"], atol=1e-4)


@pytest.mark.parametrize("name,cfg_name", [("gelu", "TEXT_TINY"), ("quick_gelu", "TEXT_TINY_QUICK")])
def test_text_encoder_golden_from_transformers(name, cfg_name):
    from ddpo_b200 import text_encoder as T
    from oracle import text_encoder as OT
    z = np.load(os.path.join(G, f"text_tiny_{name}.npz"))
    cfg = getattr(T, cfg_name)
    flat = T.init_flat_params(cfg, 0)
    assert abs(flat.double().sum().item() - float(z["param_sum"])) < 1e-6
    got = OT.encode(T.views(flat, cfg), cfg, z["input_ids"])
    np.testing.assert_allclose(got.numpy(), z["last_hidden_state"], rtol=2e-4, atol=2e-5)
