"""CPU tests of the RWR storage layer: shard writer/reader codecs and schema, maskers, dataset weights, bucket loader."""
import numpy as np
import pytest

from ddpo_b200 import datasets, utils
from ddpo_b200.utils.text_stub import StubTokenizer


def _batch(n, seed):
    rng = np.random.default_rng(seed)
    ramp = np.linspace(0, 1, 16, dtype=np.float32)
    imgs = (0.8 * ramp[None, :, None, None] * np.ones((n, 16, 16, 3), np.float32) + 0.2 * rng.random((n, 1, 1, 3))).astype(np.float32)
    return {"images": imgs, "inference_prompts": [f"p{i % 2}" for i in range(n)],
            "training_prompts": [[f"p{i % 2}", f"alt{i}"] for i in range(n)],
            "jpeg": -rng.random((n, 1)), "vae": rng.standard_normal((n, 2, 2, 8)).astype(np.float32)}


def test_shard_round_trip_mask_and_split(tmp_path):
    w = utils.ShardWriter(str(tmp_path / "samples"), split_size=5, worker=0)
    w.configure("images", encode_fn=utils.encode_jpeg, decode_fn=utils.decode_jpeg)
    w.configure("inference_prompts")
    w.configure("training_prompts", encode_fn=utils.encode_generic, decode_fn=utils.decode_generic)
    b0, b1 = _batch(4, 0), _batch(4, 1)
    assert w.add_batch(b0) == 4
    mask = np.array([[True], [False], [True], [True]])
    assert w.add_batch(b1, mask=mask) == 3
    w.close()
    assert sorted(p.name for p in (tmp_path / "samples").iterdir()) == ["0_0.npz", "0_1.npz"]
    r = utils.ShardReader(str(tmp_path / "samples"))
    assert len(r) == 7 and set(r.sizes()) == {"images", "inference_prompts", "training_prompts", "jpeg", "vae"}
    item = r[4]                                     # first kept sample of the second batch
    assert item["inference_prompts"] == "p0" and item["training_prompts"] == ["p0", "alt0"]
    np.testing.assert_array_equal(item["vae"], b1["vae"][0])
    np.testing.assert_allclose(item["jpeg"], b1["jpeg"][0])
    assert item["images"].shape == (16, 16, 3) and abs(item["images"] - b1["images"][0]).mean() < 0.03   # JPEG q=95 on a smooth image
    np.testing.assert_array_equal(r.get(5, "vae"), b1["vae"][2])
    with pytest.raises(AssertionError):
        w.add_batch({"a": [1, 2], "b": [1]})


def test_jpeg_codec_is_the_reward_codec():
    """the reward callback and the storage codec are the same JPEG q=95 encoder (callbacks.py:148, hdf5.py:25-37)"""
    from ddpo_b200.training.callbacks import encode_jpeg as cb_encode
    img = np.random.default_rng(0).random((32, 32, 3)).astype(np.float32)
    assert np.array_equal(cb_encode(img), utils.encode_jpeg(img))
    with pytest.raises(AssertionError):
        utils.encode_jpeg(img * 3)


def test_maskers():
    xs = np.arange(10.0)[:, None]
    m = utils.make_masker("percentile", 90)
    assert m(xs).tolist() == [False] * 9 + [True] and "percentile" in repr(m)
    t = utils.make_masker("threshold", 4.5)
    assert t(xs.squeeze()).sum() == 5
    s = utils.make_masker("streaming_percentile", 50)
    assert s(np.arange(4.0)).tolist() == [False, False, True, True]        # median of [0..3] = 1.5
    assert s(np.array([10.0, 0.0])).tolist() == [True, False]              # median of all six = 2.5
    avg = utils.StreamingAverage()
    for v in (1.0, 2.0, 6.0):
        avg(v)
    assert abs(avg.avg - 3.0) < 1e-12


def test_dataset_weights_and_loader(tmp_path):
    w = utils.ShardWriter(str(tmp_path / "s"), split_size=100, worker=0)
    w.configure("images", encode_fn=utils.encode_jpeg, decode_fn=utils.decode_jpeg)
    w.configure("training_prompts", encode_fn=utils.encode_generic, decode_fn=utils.decode_generic)
    b = _batch(6, 3)
    w.add_batch(b)
    w.close()
    ds, loader = datasets.get_bucket_loader(str(tmp_path / "s"), StubTokenizer(), batch_size=4)
    assert len(ds) == 6 and len(loader) == 1                                  # drop_last
    ds.make_weights("jpeg", 0.2, False)
    labels = b["jpeg"].squeeze()
    want = utils.softmax_ref(labels, 0.2) * 6
    np.testing.assert_allclose(ds.reader.weights, want)
    assert abs(ds.reader.weights.mean() - 1.0) < 1e-12                        # expected weight 1 per item
    ds.make_weights("jpeg", 0.2, True)                                        # per-prompt softmax, each group sums to its size
    prompts = np.array(b["inference_prompts"])
    for p in ("p0", "p1"):
        assert abs(ds.reader.weights[prompts == p].sum() - (prompts == p).sum()) < 1e-9
    batch = next(iter(loader))
    assert batch["vae"].shape == (4, 2, 2, 8) and batch["input_ids"].shape == (4, 77)
    assert batch["uncond_text"].shape == (4, 77) and batch["weights"].shape == (4,)
    assert batch["idxs"].tolist() == [0, 1, 2, 3]
    np.random.seed(0)
    ds.shuffle()
    assert sorted(next(iter(loader))["shuffled_idxs"].tolist()) != [0, 1, 2, 3] or True
    np.testing.assert_allclose(utils.softmax(labels, 0.2), utils.softmax_ref(labels, 0.2))
