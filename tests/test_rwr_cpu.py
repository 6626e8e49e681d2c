"""CPU tests of the RWR additions: oracle restatement of ``ddpo/training/diffusion.py`` (self-consistency pins -- the
reference ships no fixtures and JAX is not installable offline), and the C-ABI host ``randint`` against the oracle."""
import numpy as np
import torch

from oracle import diffusion as OD, scheduler as OS, threefry


def test_randint_range_determinism_and_host_abi_agreement():
    from ddpo_b200 import ops
    for seed in (0, 1, 2 ** 40 + 3):
        key = threefry.PRNGKey(seed)
        for n in (1, 2, 5, 64):
            r = threefry.randint(key, (n,), 0, 1000)
            assert r.dtype == np.int32 and r.min() >= 0 and r.max() < 1000
            assert np.array_equal(r, threefry.randint(key, (n,), 0, 1000))
            assert ops.threefry_randint((int(key[0]), int(key[1])), n, 0, 1000) == list(r)
    # span 1 -> always minval; a power-of-two span reduces to the low bits of `lower` plus a multiple of 0
    assert np.all(threefry.randint(threefry.PRNGKey(9), (7,), 5, 6) == 5)
    big = threefry.randint(threefry.PRNGKey(4), (20000,), 0, 1000)
    assert abs(big.mean() - 499.5) < 10 and len(np.unique(big)) == 1000


def test_randint_equals_wide_integer_formula():
    """the uint32 two-draw construction equals ((higher * 2**32 + lower) mod span) computed in Python integers"""
    key = threefry.PRNGKey(77)
    k1, k2 = threefry.split(key)
    hi = threefry.random_bits(k1, (33,)).astype(object)
    lo = threefry.random_bits(k2, (33,)).astype(object)
    for span in (1000, 7, 65536, 3):
        want = np.array([(int(h) * 2 ** 32 + int(l)) % span for h, l in zip(hi, lo)], np.int64)
        got = threefry.randint(key, (33,), 0, span)
        assert np.array_equal(got, want)


def test_make_inputs_identities():
    ac = OS.create_state(OS.SD_CONFIG).alphas_cumprod
    g = torch.Generator().manual_seed(0)
    mom = torch.randn(3, 8, 8, 8, generator=g).numpy()
    _, srng, new_rng = OD.split3(threefry.PRNGKey(1))
    noisy, noise, ts, lat = OD.make_inputs(mom, srng, ac)
    assert noisy.shape == noise.shape == lat.shape == (3, 4, 8, 8) and ts.shape == (3,)
    # noise is the plain NCHW normal of the second split; timesteps from the third
    nr, tr = threefry.split(srng)
    assert np.array_equal(noise, threefry.normal(nr, (3, 4, 8, 8)))
    assert np.array_equal(ts, threefry.randint(tr, (3,), 0, 1000))
    # add_noise identity and the posterior sample with logvar -> -inf (clipped at -30) collapsing onto the mean
    sa = np.sqrt(ac[ts]).reshape(-1, 1, 1, 1)
    sb = np.sqrt(1 - ac[ts]).reshape(-1, 1, 1, 1)
    np.testing.assert_allclose(noisy, sa * lat + sb * noise, rtol=1e-6, atol=1e-7)
    mom2 = mom.copy()
    mom2[..., 4:] = -1e9
    _, _, _, lat2 = OD.make_inputs(mom2, srng, ac)
    np.testing.assert_allclose(lat2, np.transpose(mom2[..., :4], (0, 3, 1, 2)) * 0.18215, rtol=1e-5, atol=1e-6)
    assert not np.array_equal(new_rng, srng)


def test_mse_loss_weighting_and_cfg():
    g = torch.Generator().manual_seed(2)
    eu, ec, nz = (torch.randn(4, 64, generator=g) for _ in range(3))
    l, per = OD.mse_loss(eu, ec, nz, 5.0, True, None)
    pred = eu + 5.0 * (ec - eu)
    np.testing.assert_allclose(per.numpy(), ((nz - pred) ** 2).mean(1).numpy(), rtol=1e-6)
    np.testing.assert_allclose(l.item(), per.mean().item(), rtol=1e-6)
    w = torch.tensor([0.1, 0.2, 0.3, 0.4])
    lw, _ = OD.mse_loss(eu, ec, nz, 5.0, True, w)
    np.testing.assert_allclose(lw.item(), (per * w).sum().item(), rtol=1e-6)
    # uniform weights 1/B reproduce the mean (pipeline/finetune.py:176-178: weights / pod_batch_size)
    lu, _ = OD.mse_loss(eu, ec, nz, 5.0, True, torch.full((4,), 0.25))
    np.testing.assert_allclose(lu.item(), l.item(), rtol=1e-6)
    l0, _ = OD.mse_loss(eu, ec, nz, 5.0, False, None)
    np.testing.assert_allclose(l0.item(), ((nz - ec) ** 2).mean().item(), rtol=1e-6)
