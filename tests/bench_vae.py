"""Timing helper (not a test): VAE decode of 8 SD-size latents and CLIP text encode of 8 prompts, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddpo_b200 import ops  # noqa: E402
from ddpo_b200.text_encoder import SD2_TEXT, CLIPTextEncoder  # noqa: E402
from ddpo_b200.vae import SD_VAE, VAEDecoder  # noqa: E402

dev = "cuda"
dec = VAEDecoder(SD_VAE, device=dev, seed=0, decode_batch=2)
enc = CLIPTextEncoder(SD2_TEXT, device=dev, seed=1)
lat = (torch.randn(8, 4, 64, 64) * 0.18215).to(dev)
ids = torch.randint(3, 49408, (8, 77)).numpy()
for name, fn in (("vae_decode_8", lambda: dec.decode(lat)), ("text_encode_8", lambda: enc(ids))):
    fn()
    torch.cuda.synchronize()
    ops.PROFILE, ops.PROFILE_TAGS = [], []
    fn()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for n, work, a, b in prof:
        d = agg.setdefault(n, [0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(name, "ms", round(e0.elapsed_time(e1) / 3, 3), {k: (round(v[0], 3), v[1]) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])})
