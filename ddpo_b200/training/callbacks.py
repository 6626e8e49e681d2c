"""Reward callbacks -- mirrors the registry/signature of ``ddpo/training/callbacks.py:540-564``:
``callback_fns[name]() -> fn(images f32 NHWC in [0,1] [N,H,W,3], prompts, metadata) -> (scores, info)``.
``jpeg`` / ``neg_jpeg`` (:143-163, JPEG q=95 size in kB via PIL, ``hdf5.py:25-37``) and ``arange`` (:347-354)
are real; the model-based rewards (aesthetic CLIP+MLP, LLaVA over HTTP) need weights / a server that do not
exist offline and are provided as cached-score stubs (BASELINE.json configs 3 and 5)."""
import io
import random

import numpy as np


def encode_jpeg(x, quality=95):
    from PIL import Image
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating):
        assert np.abs(x).max() <= 1.0
        x = (x * 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(x).save(buf, "JPEG", quality=quality)
    return np.frombuffer(buf.getvalue(), dtype=np.uint8)


def jpeg_fn(devices=None, jit=False):
    assert not jit

    def _fn(images, prompts, metadata):
        sizes_kb = [len(encode_jpeg(im)) / 1000.0 for im in images]
        return -np.array(sizes_kb)[:, None], {}
    return _fn


def neg_jpeg_fn(*args, **kwargs):
    inner = jpeg_fn(*args, **kwargs)

    def _fn(*a, **k):
        scores, info = inner(*a, **k)
        return -scores, info
    return _fn


def arange_fn(devices=None, jit=False):
    def _fn(images, prompts, metadata):
        return np.arange(len(images)), {}
    return _fn


def _cached_score_stub(name, shape_2d):
    def factory(devices=None, jit=False):
        def _fn(images, prompts, metadata):
            rng = np.random.default_rng(abs(hash((name,) + tuple(prompts))) % (2 ** 32))
            s = rng.standard_normal(len(images)).astype(np.float32)
            return (s[:, None] if shape_2d else s), {"stub": True}
        return _fn
    return factory


def evaluate_callbacks(fns, images, prompts, metadata):
    if type(prompts[0]) == list:
        prompts = [random.choice(p) for p in prompts]
    images = np.asarray(images, dtype=np.float32)
    return {key: fn(images, prompts, metadata) for key, fn in fns.items()}


callback_fns = {
    "jpeg": jpeg_fn,
    "neg_jpeg": neg_jpeg_fn,
    "arange": arange_fn,
    "aesthetic": _cached_score_stub("aesthetic", True),
    "llava_bertscore": _cached_score_stub("llava_bertscore", False),
}
