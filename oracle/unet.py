"""torch-CPU restatement of ``FlaxUNet2DConditionModel.__call__`` (diffusers==0.12.1:
``models/unet_2d_condition_flax.py``, ``unet_2d_blocks_flax.py``, ``resnet_flax.py``,
``attention_flax.py``, ``embeddings_flax.py`` -- third-party, un-vendored; the
reference calls it at ``pipeline_flax_stable_diffusion.py:219-224`` and
``training/policy_gradient.py:87-102``).

Flax semantics restated here (they differ from the PyTorch SD model):
NHWC activations / HWIO kernels; GroupNorm eps 1e-5 with var = max(0, E[x^2]-E[x]^2);
tanh-GELU; attention scale applied after QK^T; ``attention_head_dim`` = number of
heads; stride-2 down conv with symmetric pad 1; nearest x2 up-sample = in[i//2].

Differentiable (torch autograd) so the same code gives reference gradients.
``dtype`` float32 (the reference's) or float64 (error-budget truth).
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F


def _conv(x, p, name, stride=1, pad=1):
    # x NHWC, kernel HWIO
    # contiguous OIHW weight: the permuted view sends oneDNN down a ~10x slower forward path (same values)
    w = p[name + "/kernel"].permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(x.permute(0, 3, 1, 2), w, p[name + "/bias"], stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1)


def _dense(x, p, name, bias=True):
    y = x @ p[name + "/kernel"]
    return y + p[name + "/bias"] if bias else y


def group_norm(x, scale, bias, groups=32, eps=1e-5):
    b, h, w, c = x.shape
    xg = x.reshape(b, h * w, groups, c // groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    mean2 = (xg * xg).mean(dim=(1, 3), keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    y = (xg - mean) * torch.rsqrt(var + eps)
    return y.reshape(b, h, w, c) * scale + bias


def layer_norm(x, scale, bias, eps=1e-5):
    mean = x.mean(dim=-1, keepdim=True)
    mean2 = (x * x).mean(dim=-1, keepdim=True)
    var = torch.clamp(mean2 - mean * mean, min=0.0)
    return (x - mean) * torch.rsqrt(var + eps) * scale + bias


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def silu(x):
    return x * torch.sigmoid(x)


def timestep_embedding(t, dim, dtype):
    half = dim // 2
    inc = math.log(10000.0) / half  # freq_shift = 0
    inv = torch.exp(torch.arange(half, dtype=torch.float32) * -inc)
    emb = t.to(torch.float32)[:, None] * inv[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=1).to(dtype)  # flip_sin_to_cos


def _attention(x, ctx, p, name, heads):
    b, n, c = x.shape
    d = c // heads
    q = _dense(x, p, name + "/to_q", False)
    k = _dense(ctx, p, name + "/to_k", False)
    v = _dense(ctx, p, name + "/to_v", False)
    split = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)
    q, k, v = split(q), split(k), split(v)
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    o = torch.softmax(s, dim=-1) @ v
    o = o.permute(0, 2, 1, 3).reshape(b, n, c)
    return _dense(o, p, name + "/to_out_0")


class UNetOracle:
    def __init__(self, cfg, params: Dict[str, torch.Tensor], dtype=torch.float32,
                 tap: Optional[Callable[[str, torch.Tensor], None]] = None):
        self.cfg = cfg
        self.p = {k: v.to(dtype) for k, v in params.items()}
        self.dtype = dtype
        self.tap = tap or (lambda name, t: None)

    def resnet(self, x, temb_act, name):
        p = self.p
        h = silu(group_norm(x, p[name + "/norm1/scale"], p[name + "/norm1/bias"]))
        h = _conv(h, p, name + "/conv1")
        h = h + _dense(temb_act, p, name + "/time_emb_proj")[:, None, None, :]
        h = silu(group_norm(h, p[name + "/norm2/scale"], p[name + "/norm2/bias"]))
        h = _conv(h, p, name + "/conv2")
        if name + "/conv_shortcut/kernel" in p:
            x = _conv(x, p, name + "/conv_shortcut", pad=0)
        out = h + x
        self.tap(name, out)
        return out

    def transformer(self, x, ctx, name, heads):
        p = self.p
        b, hh, ww, c = x.shape
        res = x
        h = group_norm(x, p[name + "/norm/scale"], p[name + "/norm/bias"])
        if self.cfg.use_linear_projection:
            h = _dense(h.reshape(b, hh * ww, c), p, name + "/proj_in")
        else:
            h = _conv(h, p, name + "/proj_in", pad=0).reshape(b, hh * ww, c)
        bl = name + "/transformer_blocks_0"
        h = h + _attention(layer_norm(h, p[bl + "/norm1/scale"], p[bl + "/norm1/bias"]),
                           layer_norm(h, p[bl + "/norm1/scale"], p[bl + "/norm1/bias"]), p, bl + "/attn1", heads)
        self.tap(bl + "/attn1", h)
        h = h + _attention(layer_norm(h, p[bl + "/norm2/scale"], p[bl + "/norm2/bias"]), ctx, p, bl + "/attn2", heads)
        self.tap(bl + "/attn2", h)
        f = _dense(layer_norm(h, p[bl + "/norm3/scale"], p[bl + "/norm3/bias"]), p, bl + "/ff/net_0/proj")
        lin, gate = f.chunk(2, dim=-1)
        h = h + _dense(lin * gelu_tanh(gate), p, bl + "/ff/net_2")
        self.tap(bl + "/ff", h)
        if self.cfg.use_linear_projection:
            h = _dense(h, p, name + "/proj_out").reshape(b, hh, ww, c)
        else:
            h = _conv(h.reshape(b, hh, ww, c), p, name + "/proj_out", pad=0)
        out = h + res
        self.tap(name, out)
        return out

    def __call__(self, sample, timesteps, ctx):
        """sample [B,4,H,W] (NCHW, as the reference passes it), timesteps [B] int, ctx [B,L,D]."""
        cfg, p = self.cfg, self.p
        dt = self.dtype
        boc = cfg.block_out_channels
        x = sample.to(dt).permute(0, 2, 3, 1)
        ctx = ctx.to(dt)
        temb = timestep_embedding(torch.as_tensor(timesteps), boc[0], dt)
        temb = _dense(silu(_dense(temb, p, "time_embedding/linear_1")), p, "time_embedding/linear_2")
        self.tap("temb", temb)
        ta = silu(temb)
        x = _conv(x, p, "conv_in")
        self.tap("conv_in", x)
        skips = [x]
        for i, c in enumerate(boc):
            for l in range(cfg.layers_per_block):
                x = self.resnet(x, ta, f"down_blocks_{i}/resnets_{l}")
                if cfg.down_has_attn[i]:
                    x = self.transformer(x, ctx, f"down_blocks_{i}/attentions_{l}", cfg.attention_head_dim[i])
                skips.append(x)
            if i < len(boc) - 1:
                x = _conv(x, p, f"down_blocks_{i}/downsamplers_0/conv", stride=2, pad=1)
                self.tap(f"down_blocks_{i}/downsamplers_0", x)
                skips.append(x)
        x = self.resnet(x, ta, "mid_block/resnets_0")
        x = self.transformer(x, ctx, "mid_block/attentions_0", cfg.attention_head_dim[-1])
        x = self.resnet(x, ta, "mid_block/resnets_1")
        rev_heads = tuple(reversed(cfg.attention_head_dim))
        has_attn = tuple(reversed(cfg.down_has_attn))
        for i in range(len(boc)):
            for l in range(cfg.layers_per_block + 1):
                x = torch.cat([x, skips.pop()], dim=-1)
                x = self.resnet(x, ta, f"up_blocks_{i}/resnets_{l}")
                if has_attn[i]:
                    x = self.transformer(x, ctx, f"up_blocks_{i}/attentions_{l}", rev_heads[i])
            if i < len(boc) - 1:
                x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)  # nearest x2: in[i//2]
                x = _conv(x, p, f"up_blocks_{i}/upsamplers_0/conv")
                self.tap(f"up_blocks_{i}/upsamplers_0", x)
        x = silu(group_norm(x, p["conv_norm_out/scale"], p["conv_norm_out/bias"]))
        x = _conv(x, p, "conv_out")
        return x.permute(0, 3, 1, 2)
