"""CPU tests of the epoch driver's host logic (ddpo_b200/pipeline/policy_gradient.py) against a literal NumPy
restatement of the reference's procedure (pipeline/policy_gradient.py:375-441), plus serialization and text stubs."""
import os

import numpy as np
import pytest
import torch

from ddpo_b200.pipeline import policy_gradient as PG


def _reference_minibatches(N, T, B, num_train_ts, accumulation, seed):
    """what the reference feeds p_train_step, as (sample, original step) codes, n_devices = 1"""
    np.random.seed(seed)
    code = np.arange(N)[:, None] * 1000 + np.arange(T)[None, :]
    samples = {"latents": code, "advantages": np.arange(N) * 10}
    perm = np.random.permutation(N)                                              # :385
    samples = {k: v[perm] for k, v in samples.items()}
    perms = np.array([np.random.permutation(T) for _ in range(N)])               # :390
    samples["latents"] = samples["latents"][np.arange(N)[:, None], perms]       # :394
    train = {k: v.reshape(-1, 1, B, *v.shape[1:]) for k, v in samples.items()}   # :398
    out = []
    n_mb = train["latents"].shape[0]
    for i in range(n_mb):
        for j in range(num_train_ts):
            upd = (j == num_train_ts - 1) and ((i + 1) % accumulation == 0)      # :426
            out.append((train["latents"][i, 0, :, j].tolist(), train["advantages"][i, 0].tolist(), upd))
    return out


@pytest.mark.parametrize("N,T,B,ratio,acc,macro", [(8, 6, 2, 1.0, 1, 1), (8, 6, 2, 1.0, 2, 3), (12, 10, 4, 0.5, 1, 5),
                                                   (4, 50, 2, 1.0, 2, 10), (6, 7, 3, 1.0, 1, 10)])
def test_train_schedule_equals_reference_procedure(N, T, B, ratio, acc, macro):
    num_train_ts = int(T * ratio)
    ref = _reference_minibatches(N, T, B, num_train_ts, acc, seed=5)
    np.random.seed(5)
    perm, perms = PG.epoch_shuffles(N, T)
    J = PG.pick_macro(num_train_ts, macro)
    assert num_train_ts % J == 0 and J <= max(1, macro)
    got = []
    adv = np.arange(N) * 10
    for step in PG.train_schedule(perm, perms, B, num_train_ts, acc, J):
        codes = (step["sample_idx"] * 1000 + step["time_idx"]).reshape(J, B)
        advs = adv[step["sample_idx"]].reshape(J, B)
        for jj in range(J):
            last = step["do_opt_update"] and jj == J - 1
            got.append((codes[jj].tolist(), advs[jj].tolist(), last))
    assert got == ref
    assert got[-1][2]  # the loop ends on an optimizer update (:446)


def test_batch_sizes_and_asserts():
    args = PG.Parser().parse_args("pg", ["--sample_batch_size", "8", "--num_sample_batches_per_epoch", "2"])
    s = PG.batch_sizes(args, 1, 4)
    assert s["total_samples_per_epoch"] == 64 and s["train_effective_batch_size"] == 8
    assert s["updates_per_inner_epoch"] == 8
    args.train_batch_size = 3
    with pytest.raises(AssertionError):
        PG.batch_sizes(args, 1, 4)


def test_compute_advantages_matches_reference_formulas():
    from ddpo_b200.utils.stat_tracking import PerPromptStatTracker
    r = np.array([1.0, 2.0, 4.0, 8.0])
    np.testing.assert_allclose(PG.compute_advantages(r, ["a"] * 4, None), (r - r.mean()) / r.std())
    tr = PerPromptStatTracker(32, 2)
    a = PG.compute_advantages(r, np.array(["x", "x", "y", "y"]), tr)
    np.testing.assert_allclose(a[:2], (r[:2] - 1.5) / (0.5 + 1e-6))
    np.testing.assert_allclose(a[2:], (r[2:] - 6.0) / (2.0 + 1e-6))


def test_seed_lineage_matches_oracle_prng():
    """PRNGKey(seed) -> split -> (train_rng, sample_rng) -> per batch split -> split(n_devices=1)[0] (:51,201,244-245)"""
    from ddpo_b200 import ops
    from oracle import threefry
    rng = ops.prng_key(42)
    train_rng, sample_rng = ops.threefry_split(rng, 2)
    o_train, o_sample = threefry.split(threefry.PRNGKey(42))
    assert tuple(int(v) for v in o_sample) == sample_rng and tuple(int(v) for v in o_train) == train_rng
    sample_rng, sample_seed = ops.threefry_split(sample_rng, 2)
    seeds = ops.threefry_split(sample_seed, 1)
    o_rng, o_seed = threefry.split(o_sample)
    assert tuple(int(v) for v in threefry.split(o_seed, 1)[0]) == seeds[0]


def test_flax_msgpack_and_pkl_checkpoints_round_trip(tmp_path):
    from ddpo_b200 import unet_spec, utils
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 3)
    tree = utils.params_tree(flat, cfg)
    assert utils.n_params(tree) == unet_spec.num_params(cfg)
    assert tree["down_blocks_0"]["resnets_0"]["conv1"]["kernel"].shape == (3, 3, 64, 64)
    path = utils.save_checkpoint(str(tmp_path / "checkpoints"), tree, step=7, keep=2, overwrite=True)
    assert os.path.basename(path) == "checkpoint_7"
    utils.save_checkpoint(str(tmp_path / "checkpoints"), tree, step=8, keep=2)
    utils.save_checkpoint(str(tmp_path / "checkpoints"), tree, step=9, keep=2)
    assert sorted(os.listdir(tmp_path / "checkpoints")) == ["checkpoint_8", "checkpoint_9"]
    with pytest.raises(FileExistsError):
        utils.save_checkpoint(str(tmp_path / "checkpoints"), tree, step=9, keep=2)
    back = utils.restore_checkpoint(str(tmp_path / "checkpoints"))
    assert torch.equal(utils.flat_from_tree(back, cfg), flat)
    # the msgpack layout: ndarray = ExtType(1, packb((shape, dtype.name, bytes)))
    import msgpack
    raw = msgpack.unpackb(open(tmp_path / "checkpoints" / "checkpoint_9", "rb").read(), raw=False, strict_map_key=False)
    ext = raw["conv_in"]["bias"]
    assert isinstance(ext, msgpack.ExtType) and ext.code == 1
    shape, dtype, buf = msgpack.unpackb(ext.data, raw=False)
    assert shape == [64] and dtype == "float32" and len(buf) == 256
    # unet_<epoch>.pkl (RWR handoff, reference serialization.py:276-317)
    utils.save_unet(str(tmp_path / "models"), tree, epoch=3)
    utils.save_unet(str(tmp_path / "models"), tree, epoch=10)
    assert utils.get_latest_epoch(str(tmp_path / "models" / "unet")) == 10
    tree2 = utils.load_flax_model(str(tmp_path / "models" / "unet"), epoch="latest")
    assert torch.equal(utils.flat_from_tree(tree2, cfg), flat)


def test_text_stubs_are_deterministic_and_round_trip():
    from ddpo_b200.utils.text_stub import StubTextEncoder, StubTokenizer
    tok = StubTokenizer()
    ids = tok(["a dog", "A  dog", "a zebra riding a bike", ""], padding="max_length", return_tensors="np").input_ids
    assert ids.shape == (4, 77) and np.array_equal(ids[0], ids[1]) and not np.array_equal(ids[0], ids[2])
    assert tok.batch_decode(ids, skip_special_tokens=True) == ["a dog", "a dog", "a zebra riding a bike", ""]
    # decoding needs no state: a tokenizer that never saw these prompts (another rank) gives the same strings
    assert StubTokenizer().batch_decode(ids) == ["a dog", "a dog", "a zebra riding a bike", ""]
    enc = StubTextEncoder(64)
    e = enc(ids)[0]
    assert e.shape == (4, 77, 64) and e.dtype == np.float32
    assert np.array_equal(e[0], e[1]) and not np.array_equal(e[0], e[2])
    assert abs(float(e.std()) - 1.0) < 0.05
    assert np.array_equal(enc(ids)[0], e)
