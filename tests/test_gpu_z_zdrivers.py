"""End-to-end runs of the three drivers (ddpo_b200/pipeline/{policy_gradient,sample,finetune}.py) on a tiny random-init
model: DDPO epoch loop (sample -> VAE decode -> jpeg reward -> advantages -> shuffled PPO updates -> checkpoints) and
the RWR loop (sample -> shards -> weighted fine-tune).  Also the on-device trajectory gather against direct indexing."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models(seed=0, text_encoder="stub"):
    from ddpo_b200 import utils
    return utils.load_unet(None, pretrained_model="tiny", device="cuda", seed=seed, text_encoder=text_encoder)


def test_epoch_buffer_gather_equals_direct_indexing():
    from ddpo_b200.pipeline import policy_gradient as PG
    dev = "cuda"
    N, T, b = 6, 5, 3
    g = torch.Generator().manual_seed(0)
    buf = PG.EpochBuffer(N, T, (4, 8, 8), (7, 16), dev)
    lat_all = torch.randn(N, T, 4, 8, 8, generator=g)
    fin_all = torch.randn(N, 4, 8, 8, generator=g)
    lps = torch.randn(N, T, generator=g)
    ts = torch.arange(T, 0, -1, dtype=torch.int32)[None].expand(N, T)
    emb = torch.randn(N, 7, 16, generator=g)
    for s in range(0, N, b):
        buf.append(fin_all[s:s + b].to(dev), lat_all[s:s + b].to(dev), lps[s:s + b].to(dev), ts[s:s + b].to(dev),
                   emb[s:s + b].to(dev))
    nxt_all = torch.cat([lat_all[:, 1:], fin_all[:, None]], dim=1)
    np.random.seed(1)
    perm, perms = PG.epoch_shuffles(N, T)
    adv = np.linspace(-1, 1, N).astype(np.float32)
    unc = torch.zeros(1, 7, 16, device=dev)
    for step in PG.train_schedule(perm, perms, 2, T, 1, 1):
        batch = buf.gather(step["sample_idx"], step["time_idx"], adv, unc)
        si, ti = step["sample_idx"], step["time_idx"]
        assert torch.equal(batch["latents"].cpu(), lat_all[si, ti])
        assert torch.equal(batch["next_latents"].cpu(), nxt_all[si, ti])
        assert torch.equal(batch["prompt_embeds"].cpu(), emb[si])
        assert torch.equal(batch["log_probs"].cpu(), lps[si, ti])
        assert batch["ts"].cpu().tolist() == ts[si, ti].tolist()
        np.testing.assert_array_equal(batch["advantages"].cpu().numpy(), adv[si])


def test_ddpo_driver_two_epochs(tmp_path, monkeypatch):
    from ddpo_b200 import unet_spec, utils
    from ddpo_b200.pipeline import policy_gradient as PG
    from ddpo_b200.training import policy_gradient as pg
    monkeypatch.chdir(tmp_path)
    pg._GRAPHS.clear()
    models = _models()
    p0 = models[1]["unet"].clone()
    argv = ["--dataset", "compressed_animals", "--pretrained_model", "tiny", "--resolution", "128",
            "--sample_batch_size", "4", "--num_sample_batches_per_epoch", "2", "--n_inference_steps", "4",
            "--train_batch_size", "2", "--train_macro", "2", "--num_train_epochs", "2", "--save_freq", "1",
            "--learning_rate", "1e-4", "--savepath", "run0", "--seed", "3"]
    out = PG.main(argv, models=models, max_epochs=2)
    hist = out["history"]
    assert len(hist) == 2 and all(np.isfinite(h["mean_reward"]) for h in hist)
    info = hist[0]["infos"][0]
    # 8 samples / train batch 2 = 4 minibatches x (4 timesteps / macro 2) = 8 passes, one optimizer update per minibatch
    assert info["approx_kl"].shape == (8,)
    # policy unchanged until the first update: ratio == 1 exactly for the first minibatch's passes
    assert info["approx_kl"][0] == 0.0 and info["approx_kl"][1] == 0.0 and info["clipfrac"][0] == 0.0
    assert info["approx_kl"][2:].max() > 0.0                       # after the first AdamW step the policy moved
    assert out["state"].step == 8 and not torch.equal(models[1]["unet"], p0)
    lp = out["localpath"]
    for rel in ("args.json", "samples/0_0_0.png", "samples/0_1_1.png", "rewards/0_1.npy", "prompts/0_0.npy",
                "callback_info/0_0.npy", "per_prompt_stats/0_1.npy", "train_info/0_1_0.npy"):
        assert os.path.exists(os.path.join(lp, rel)), rel
    r = np.load(os.path.join(lp, "rewards/0_0.npy"))
    assert r.shape == (8, 1) and (r < 0).all()                      # jpeg reward = -kB
    tree = utils.restore_checkpoint(os.path.join("logs/compressed-animals/run0", "checkpoints"))
    assert torch.equal(utils.flat_from_tree(tree, unet_spec.TINY), out["state"].params.cpu())


def test_ddpo_driver_macro_equals_reference_call_sequence(tmp_path, monkeypatch):
    """train_macro = 1 (the reference's one train_step per timestep) and train_macro = 4 give the same parameters
    after an epoch up to fp32 summation order"""
    from ddpo_b200.pipeline import policy_gradient as PG
    from ddpo_b200.training import policy_gradient as pg
    monkeypatch.chdir(tmp_path)
    res = []
    for macro in ("1", "4"):
        pg._GRAPHS.clear()
        models = _models(seed=1)
        argv = ["--dataset", "compressed_animals", "--pretrained_model", "tiny", "--resolution", "128",
                "--sample_batch_size", "2", "--num_sample_batches_per_epoch", "1", "--n_inference_steps", "4",
                "--train_batch_size", "2", "--train_macro", macro, "--num_train_epochs", "1", "--save_freq", "100",
                "--learning_rate", "1e-4", "--savepath", f"m{macro}", "--seed", "5", "--per_prompt_stats_bufsize", "None",
                "--filter_field", "arange"]
        out = PG.main(argv, models=models, max_epochs=1)
        res.append((out["state"].params.clone(), out["history"][0]["infos"][0]))
    (p1, i1), (p4, i4) = res
    assert i1["approx_kl"].shape == (4,) and i4["approx_kl"].shape == (1,)
    np.testing.assert_allclose(i4["loss"][0], i1["loss"].mean(), rtol=1e-5, atol=1e-7)
    # lr 1e-4: every parameter moves ~1e-4 (Adam's first step is lr * g / (|g| + eps)); a semantic difference between
    # the two schedules would show up at that size everywhere.  Summation-order noise (1e-6 relative on g) only matters
    # for the handful of entries with |g| ~ eps = 1e-8, whose sign may flip: bounded by 2 lr.
    d = (p1 - p4).abs()
    assert d.mean().item() < 1e-6 and d.max().item() <= 2.5e-4, (d.mean().item(), d.max().item())
    assert (p1 - _models(seed=1)[1]["unet"]).abs().mean().item() > 2e-5   # and the step itself is much larger


def test_rwr_loop_sample_then_finetune(tmp_path, monkeypatch):
    from ddpo_b200 import utils
    from ddpo_b200.pipeline import finetune, sample
    from ddpo_b200.training import diffusion as D
    monkeypatch.chdir(tmp_path)
    D._GRAPHS.clear()
    models = _models(seed=2)
    common = ["--dataset", "compressed_animals_rwr", "--pretrained_model", "tiny", "--resolution", "128"]
    out = sample.main(common + ["--n_samples_per_device", "4", "--n_inference_steps", "3", "--max_steps", "2",
                                "--max_samples", "None", "--seed", "1"], models=models)
    assert out["n_steps"] == 2 and out["n_samples"] == 8          # mask_param 0: every sample is kept
    reader = utils.ShardReader(out["savepath"])
    assert len(reader) == 8 and reader[0]["vae"].shape == (16, 16, 8) and reader[0]["images"].shape == (128, 128, 3)
    p0 = models[1]["unet"].clone()
    res = finetune.main(common + ["--train_batch_size", "2", "--num_train_epochs", "2", "--save_freq", "1",
                                  "--learning_rate", "1e-4"], models=models)
    assert res["steps"] == 8 and len(res["losses"]) == 2 and all(np.isfinite(res["losses"]))
    assert not torch.equal(models[1]["unet"], p0)
    assert utils.get_latest_epoch("logs/rwr-compressed-animals/models/1/unet") == 2


def test_sampler_graph_replay_sees_each_calls_context():
    """the captured denoising step must read the K/V of the CURRENT prompt batch on every later call"""
    from ddpo_b200 import unet_spec
    from ddpo_b200.diffusers_patch import DDIMScheduler, StableDiffusionPipeline
    from ddpo_b200.unet import UNet
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 0)
    g = torch.Generator().manual_seed(7)
    embs = [torch.randn(2, cfg.ctx_len, cfg.cross_attention_dim, generator=g).cuda() for _ in range(3)]
    neg = torch.randn(1, cfg.ctx_len, cfg.cross_attention_dim, generator=g).expand(2, -1, -1).contiguous().cuda()
    outs = {}
    for use_graph in (False, True):
        net = UNet(cfg, flat, "cuda")
        sched = DDIMScheduler(1000, 0.00085, 0.012, "scaled_linear", None, False, 1, "epsilon")
        pipe = StableDiffusionPipeline(net, sched, use_cuda_graph=use_graph)
        st = sched.create_state()
        outs[use_graph] = [pipe(e, neg, {"unet": net.params, "scheduler": st}, (1, 2), 3, 128, 128, 5.0, 1.0)[0].clone()
                           for e in embs]
    torch.cuda.synchronize()
    for a, b in zip(outs[False], outs[True]):
        assert torch.equal(a, b)
    assert not torch.equal(outs[True][0], outs[True][1])


def test_ddpo_driver_with_the_gpu_text_encoder(tmp_path, monkeypatch):
    """same loop with prompts embedded by the CLIP text tower on the GPU (utils.load_unet default)"""
    from ddpo_b200.pipeline import policy_gradient as PG
    from ddpo_b200.text_encoder import CLIPTextEncoder
    from ddpo_b200.training import policy_gradient as pg
    monkeypatch.chdir(tmp_path)
    pg._GRAPHS.clear()
    models = _models(seed=4, text_encoder="clip")
    assert isinstance(models[0].text_encoder, CLIPTextEncoder)
    argv = ["--dataset", "compressed_animals", "--pretrained_model", "tiny", "--resolution", "128",
            "--sample_batch_size", "2", "--num_sample_batches_per_epoch", "1", "--n_inference_steps", "3",
            "--train_batch_size", "2", "--train_macro", "3", "--num_train_epochs", "1", "--save_freq", "100",
            "--savepath", "clip", "--seed", "1"]
    out = PG.main(argv, models=models, max_epochs=1, save_last=False)
    info = out["history"][0]["infos"][0]
    assert info["approx_kl"].shape == (1,) and info["approx_kl"][0] == 0.0 and np.isfinite(out["history"][0]["mean_reward"])
