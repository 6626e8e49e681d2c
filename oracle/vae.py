"""torch-CPU restatement of the VAE decode the reference applies to the sampled latents
(``pipeline/policy_gradient.py:174-182``; ``ddpo/training/diffusion.py:105-112``):
``latents / 0.18215`` -> 3P diffusers==0.12.1 ``FlaxAutoencoderKL.decode`` (``vae_flax.py``: ``post_quant_conv``
1x1 -> ``FlaxDecoder``: conv_in; ``FlaxUNetMidBlock2D`` = ResNet, ``FlaxAttentionBlock`` (one head,
``softmax((q s)(k s)^T) v``, ``s = (C/heads)^-1/4``), ResNet; four ``FlaxUpDecoderBlock2D`` of 3 ResNets with a
nearest-2x ``FlaxUpsample2D`` conv on all but the last; ``GroupNorm(32, eps 1e-6)`` -> swish -> conv_out)
-> ``(x / 2 + 0.5).clip(0, 1)`` NHWC.  Parameter names follow the Flax checkpoint.  TEST INFRASTRUCTURE ONLY.
Parity unpinned against a live diffusers run (not installable offline); pinned self-consistently by
``tests/test_vae_cpu.py`` (parameter count of the SD decoder, shapes, attention identity)."""
import math

import torch

from .unet import _conv, _dense, group_norm, silu

GN_EPS = 1e-6
VAE_SCALING = 0.18215


def _resnet(x, p, name):
    h = _conv(silu(group_norm(x, p[name + "/norm1/scale"], p[name + "/norm1/bias"], eps=GN_EPS)), p, name + "/conv1")
    h = _conv(silu(group_norm(h, p[name + "/norm2/scale"], p[name + "/norm2/bias"], eps=GN_EPS)), p, name + "/conv2")
    if (name + "/conv_shortcut/kernel") in p:
        x = _conv(x, p, name + "/conv_shortcut", pad=0)
    return h + x


def _attention(x, p, name):
    b, hh, ww, c = x.shape
    g = group_norm(x, p[name + "/group_norm/scale"], p[name + "/group_norm/bias"], eps=GN_EPS).reshape(b, hh * ww, c)
    q, k, v = (_dense(g, p, f"{name}/{n}") for n in ("query", "key", "value"))
    scale = 1.0 / math.sqrt(math.sqrt(c))                       # one head of width c
    w = torch.softmax((q * scale) @ (k * scale).transpose(-1, -2), dim=-1)
    o = _dense(w @ v, p, name + "/proj_attn").reshape(b, hh, ww, c)
    return o + x


def decode(params, cfg, latents, dtype=torch.float32, taps=None):
    """latents NCHW [B,4,h,w] (still scaled by 0.18215) -> (images NHWC [B,8h,8w,3] in [0,1], raw NCHW [B,3,8h,8w])."""
    p = {k: v.to(dtype) for k, v in params.items()}
    x = (torch.as_tensor(latents).to(dtype) / VAE_SCALING).permute(0, 2, 3, 1)
    x = _conv(x, p, "post_quant_conv", pad=0)
    x = _conv(x, p, "decoder/conv_in")
    x = _resnet(x, p, "decoder/mid_block/resnets_0")
    x = _attention(x, p, "decoder/mid_block/attentions_0")
    x = _resnet(x, p, "decoder/mid_block/resnets_1")
    if taps is not None:
        taps["mid"] = x
    n_up = len(cfg.block_out_channels)
    for i in range(n_up):
        for l in range(cfg.layers_per_block + 1):
            x = _resnet(x, p, f"decoder/up_blocks_{i}/resnets_{l}")
        if i < n_up - 1:
            x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)   # jax.image.resize nearest: in[i // 2]
            x = _conv(x, p, f"decoder/up_blocks_{i}/upsamplers_0/conv")
        if taps is not None:
            taps[f"up{i}"] = x
    x = silu(group_norm(x, p["decoder/conv_norm_out/scale"], p["decoder/conv_norm_out/bias"], eps=GN_EPS))
    raw = _conv(x, p, "decoder/conv_out")                        # NHWC
    images = (raw / 2 + 0.5).clamp(0, 1)
    return images, raw.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------ encoder ----
def encode(params, cfg, images_nhwc, dtype=torch.float32):
    """torch-CPU restatement of what the reference's ``vae_fn`` computes (``ddpo/training/callbacks.py:37-57``):
    ``images`` NHWC in [0, 1] -> ``(x - 0.5) / 0.5`` -> 3P diffusers==0.12.1 ``FlaxAutoencoderKL.encode`` (``vae_flax.py``:
    ``FlaxEncoder``: conv_in; per level two ResNets and -- on all but the last -- ``FlaxDownsample2D`` = pad ((0,1),(0,1)) on
    the HIGH side only, then a VALID 3x3 stride-2 conv; mid block (ResNet, single-head attention, ResNet);
    ``GroupNorm(32, eps 1e-6)`` -> swish -> conv_out to 2 x latent channels; then ``quant_conv`` 1x1) ->
    ``FlaxDiagonalGaussianDistribution``: mean, logvar = split(moments), logvar clipped to [-30, 20].
    Returns ``concatenate([mean, logvar], -1)`` NHWC ``[B, h/8, w/8, 8]``.  Parameter names follow the Flax checkpoint
    (``encoder/...``, ``quant_conv``)."""
    p = {k: v.to(dtype) for k, v in params.items()}
    x = (torch.as_tensor(images_nhwc).to(dtype) - 0.5) / 0.5
    x = _conv(x, p, "encoder/conv_in")
    n = len(cfg.block_out_channels)
    for i in range(n):
        for l in range(cfg.layers_per_block):
            x = _resnet(x, p, f"encoder/down_blocks_{i}/resnets_{l}")
        if i < n - 1:
            x = torch.nn.functional.pad(x, (0, 0, 0, 1, 0, 1))          # NHWC: pad W and H by one on the high side
            x = _conv(x, p, f"encoder/down_blocks_{i}/downsamplers_0/conv", stride=2, pad=0)
    x = _resnet(x, p, "encoder/mid_block/resnets_0")
    x = _attention(x, p, "encoder/mid_block/attentions_0")
    x = _resnet(x, p, "encoder/mid_block/resnets_1")
    x = silu(group_norm(x, p["encoder/conv_norm_out/scale"], p["encoder/conv_norm_out/bias"], eps=GN_EPS))
    x = _conv(x, p, "encoder/conv_out")
    moments = _conv(x, p, "quant_conv", pad=0)
    c = moments.shape[-1] // 2
    mean, logvar = moments[..., :c], moments[..., c:].clamp(-30.0, 20.0)
    return torch.cat([mean, logvar], dim=-1)
