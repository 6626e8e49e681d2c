"""TEST INFRASTRUCTURE: the Stable-Diffusion U-Net written the PyTorch way -- NCHW, ``torch.nn`` modules, library
primitives (``nn.GroupNorm``, ``nn.LayerNorm``, ``nn.Conv2d``, ``F.scaled_dot_product_attention``,
``F.gelu(approximate="tanh")``, ``F.interpolate(mode="nearest")``) -- with the module tree of the PyTorch
``UNet2DConditionModel`` (conv_in, time_embedding, down_blocks[i].resnets/attentions/downsamplers, mid_block, up_blocks,
conv_norm_out, conv_out).  It shares NO arithmetic with ``oracle/unet.py`` (which restates the Flax model in NHWC with
hand-written norm / attention / GELU formulas); weights are brought over with the standard Flax <-> PyTorch checkpoint
conversion (conv kernel HWIO <-> OIHW, Dense kernel [in, out] <-> Linear weight [out, in], norm scale <-> weight).
Agreement of the two pins the oracle's block arithmetic (GroupNorm eps / group partition, attention head split and
scaling, GEGLU split order, time-embedding layout, symmetric stride-2 padding, nearest up-sampling, skip wiring) to an
independent statement built from third-party primitives.  The one deliberate Flax-ism kept here is the tanh GELU
(``flax.linen.gelu`` default), which the PyTorch checkpoint code replaces by the erf form."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Timesteps(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
        emb = t.float()[:, None] * torch.exp(exponent)[None]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)      # flip_sin_to_cos=True, freq_shift=0 (fp32, as Flax)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return h + (self.conv_shortcut(x) if self.conv_shortcut is not None else x)


class Attention(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        sp = lambda t: t.view(b, t.shape[1], self.heads, c // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp(self.to_q(x)), sp(self.to_k(ctx)), sp(self.to_v(ctx)))
        return self.to_out[0](o.transpose(1, 2).reshape(b, n, c))


class GEGLU(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Linear(dim, 8 * dim)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate, approximate="tanh")


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, ctx_dim, heads):
        super().__init__()
        self.norm1, self.norm2, self.norm3 = (nn.LayerNorm(dim, eps=1e-5) for _ in range(3))
        self.attn1 = Attention(dim, dim, heads)
        self.attn2 = Attention(dim, ctx_dim, heads)
        self.ff = nn.ModuleDict({"net": nn.ModuleList([GEGLU(dim), nn.Identity(), nn.Linear(4 * dim, dim)])})

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff["net"][2](self.ff["net"][0](self.norm3(x)))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, ctx_dim, heads, linear):
        super().__init__()
        self.linear = linear
        self.norm = nn.GroupNorm(32, dim, eps=1e-5)     # Flax uses 1e-5 here (the PyTorch checkpoint code 1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, ctx_dim, heads)])
        self.proj_out = nn.Linear(dim, dim) if linear else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        y = self.norm(x)
        if self.linear:
            y = self.proj_in(y.permute(0, 2, 3, 1).reshape(b, h * w, c))
        else:
            y = self.proj_in(y).permute(0, 2, 3, 1).reshape(b, h * w, c)
        y = self.transformer_blocks[0](y, ctx)
        if self.linear:
            y = self.proj_out(y).reshape(b, h, w, c).permute(0, 3, 1, 2)
        else:
            y = self.proj_out(y.reshape(b, h, w, c).permute(0, 3, 1, 2))
        return y + res


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnets, self.attentions = nn.ModuleList(), nn.ModuleList()
        self.downsamplers, self.upsamplers = None, None


class Sampler(nn.Module):
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1)


class UNet2DConditionTwin(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        boc, te = cfg.block_out_channels, cfg.time_embed_dim
        self.cfg = cfg
        self.time_proj = Timesteps(boc[0])
        self.time_embedding = nn.ModuleDict({"linear_1": nn.Linear(boc[0], te), "linear_2": nn.Linear(te, te)})
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks, self.up_blocks = nn.ModuleList(), nn.ModuleList()
        tr = lambda c, h: Transformer2DModel(c, cfg.cross_attention_dim, h, cfg.use_linear_projection)
        prev = boc[0]
        for i, c in enumerate(boc):
            blk = Block()
            for l in range(cfg.layers_per_block):
                blk.resnets.append(ResnetBlock2D(prev if l == 0 else c, c, te))
                if cfg.down_has_attn[i]:
                    blk.attentions.append(tr(c, cfg.attention_head_dim[i]))
            if i < len(boc) - 1:
                blk.downsamplers = nn.ModuleList([Sampler(c, 2)])
            self.down_blocks.append(blk)
            prev = c
        cm = boc[-1]
        self.mid_block = Block()
        self.mid_block.resnets.extend([ResnetBlock2D(cm, cm, te), ResnetBlock2D(cm, cm, te)])
        self.mid_block.attentions.append(tr(cm, cfg.attention_head_dim[-1]))
        rev, rev_h = tuple(reversed(boc)), tuple(reversed(cfg.attention_head_dim))
        has_attn = tuple(reversed(cfg.down_has_attn))
        prev_out = rev[0]
        for i, c in enumerate(rev):
            blk = Block()
            cin = rev[min(i + 1, len(rev) - 1)]
            n = cfg.layers_per_block + 1
            for l in range(n):
                skip = cin if l == n - 1 else c
                blk.resnets.append(ResnetBlock2D((prev_out if l == 0 else c) + skip, c, te))
                if has_attn[i]:
                    blk.attentions.append(tr(c, rev_h[i]))
            if i < len(rev) - 1:
                blk.upsamplers = nn.ModuleList([Sampler(c, 1)])
            self.up_blocks.append(blk)
            prev_out = c
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timesteps, ctx):
        t_emb = self.time_proj(timesteps).to(sample.dtype)
        temb = self.time_embedding["linear_2"](F.silu(self.time_embedding["linear_1"](t_emb)))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            for l, res in enumerate(blk.resnets):
                x = res(x, temb)
                if len(blk.attentions):
                    x = blk.attentions[l](x, ctx)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0].conv(x)
                skips.append(x)
        x = self.mid_block.resnets[0](x, temb)
        x = self.mid_block.attentions[0](x, ctx)
        x = self.mid_block.resnets[1](x, temb)
        for blk in self.up_blocks:
            for l, res in enumerate(blk.resnets):
                x = res(torch.cat([x, skips.pop()], dim=1), temb)
                if len(blk.attentions):
                    x = blk.attentions[l](x, ctx)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def flax_name_to_torch(name):
    """'down_blocks_0/attentions_1/transformer_blocks_0/attn1/to_out_0/kernel' -> ('down_blocks.0.attentions.1...', kind)"""
    parts = name.split("/")
    leaf = parts[-1]
    out = []
    for p in parts[:-1]:
        head, _, idx = p.rpartition("_")
        if idx.isdigit() and head in ("down_blocks", "up_blocks", "resnets", "attentions", "transformer_blocks", "downsamplers",
                                      "upsamplers", "net"):
            out += [head, idx]
        elif p == "to_out_0":
            out += ["to_out", "0"]
        else:
            out.append(p)
    return ".".join(out), leaf


def load_flax_params(model, views):
    """Standard Flax -> PyTorch checkpoint conversion of a name -> tensor dict in Flax layout."""
    sd = model.state_dict()
    used = set()
    for name, v in views.items():
        base, leaf = flax_name_to_torch(name)
        if leaf == "kernel":
            key = base + ".weight"
            w = v.permute(3, 2, 0, 1) if v.dim() == 4 else v.t()
            if sd[key].dim() == 4 and w.dim() == 2:          # Dense checkpoint into a 1x1 conv
                w = w[:, :, None, None]
        elif leaf == "scale":
            key, w = base + ".weight", v
        else:
            key, w = base + ".bias", v
        assert key in sd, (name, key)
        assert tuple(sd[key].shape) == tuple(w.shape), (name, key, tuple(sd[key].shape), tuple(w.shape))
        sd[key] = w.clone().contiguous()
        used.add(key)
    missing = set(sd) - used
    assert not missing, sorted(missing)[:5]
    model.load_state_dict(sd)
    return model


# ------------------------------------------------------------------------------ VAE decoder twin ----
def vae_decode_twin(params, cfg, latents, scaling=0.18215):
    """The KL-VAE decoder the PyTorch way (NCHW; F.group_norm / F.conv2d / F.scaled_dot_product_attention /
    F.interpolate) on Flax-layout parameters converted on the fly.  Returns (images NHWC in [0, 1], raw NCHW)."""
    def conv(x, name, stride=1, pad=1):
        w = params[name + "/kernel"].permute(3, 2, 0, 1)
        return F.conv2d(x, w, params[name + "/bias"], stride=stride, padding=pad)

    def gn(x, name):
        return F.group_norm(x, 32, params[name + "/scale"], params[name + "/bias"], eps=1e-6)

    def lin(x, name):
        return F.linear(x, params[name + "/kernel"].t(), params[name + "/bias"])

    def resnet(x, name):
        h = conv(F.silu(gn(x, name + "/norm1")), name + "/conv1")
        h = conv(F.silu(gn(h, name + "/norm2")), name + "/conv2")
        if name + "/conv_shortcut/kernel" in params:
            x = conv(x, name + "/conv_shortcut", pad=0)
        return x + h

    def attention(x, name):
        b, c, h, w = x.shape
        g = gn(x, name + "/group_norm").permute(0, 2, 3, 1).reshape(b, 1, h * w, c)       # one head of width c
        o = F.scaled_dot_product_attention(lin(g, name + "/query"), lin(g, name + "/key"), lin(g, name + "/value"))
        return x + lin(o, name + "/proj_attn").reshape(b, h, w, c).permute(0, 3, 1, 2)

    x = conv(latents / scaling, "post_quant_conv", pad=0)
    x = conv(x, "decoder/conv_in")
    x = resnet(x, "decoder/mid_block/resnets_0")
    x = attention(x, "decoder/mid_block/attentions_0")
    x = resnet(x, "decoder/mid_block/resnets_1")
    n_up = len(cfg.block_out_channels)
    for i in range(n_up):
        for l in range(cfg.layers_per_block + 1):
            x = resnet(x, f"decoder/up_blocks_{i}/resnets_{l}")
        if i < n_up - 1:
            x = conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), f"decoder/up_blocks_{i}/upsamplers_0/conv")
    raw = conv(F.silu(gn(x, "decoder/conv_norm_out")), "decoder/conv_out")
    return (raw / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1), raw


def vae_encode_twin(params, cfg, images_nhwc):
    """The KL-VAE encoder the PyTorch way (NCHW, F.* primitives; the PyTorch Downsample2D pads (0, 1, 0, 1) before its
    padding-0 stride-2 conv) on Flax-layout parameters.  Returns moments NHWC [B, h/8, w/8, 8] (mean | clipped logvar)."""
    def conv(x, name, stride=1, pad=1):
        return F.conv2d(x, params[name + "/kernel"].permute(3, 2, 0, 1), params[name + "/bias"], stride=stride, padding=pad)

    def gn(x, name):
        return F.group_norm(x, 32, params[name + "/scale"], params[name + "/bias"], eps=1e-6)

    def lin(x, name):
        return F.linear(x, params[name + "/kernel"].t(), params[name + "/bias"])

    def resnet(x, name):
        h = conv(F.silu(gn(x, name + "/norm1")), name + "/conv1")
        h = conv(F.silu(gn(h, name + "/norm2")), name + "/conv2")
        if name + "/conv_shortcut/kernel" in params:
            x = conv(x, name + "/conv_shortcut", pad=0)
        return x + h

    def attention(x, name):
        b, c, h, w = x.shape
        g = gn(x, name + "/group_norm").permute(0, 2, 3, 1).reshape(b, 1, h * w, c)
        o = F.scaled_dot_product_attention(lin(g, name + "/query"), lin(g, name + "/key"), lin(g, name + "/value"))
        return x + lin(o, name + "/proj_attn").reshape(b, h, w, c).permute(0, 3, 1, 2)

    x = (images_nhwc.permute(0, 3, 1, 2) - 0.5) / 0.5
    x = conv(x, "encoder/conv_in")
    n = len(cfg.block_out_channels)
    for i in range(n):
        for l in range(cfg.layers_per_block):
            x = resnet(x, f"encoder/down_blocks_{i}/resnets_{l}")
        if i < n - 1:
            x = conv(F.pad(x, (0, 1, 0, 1)), f"encoder/down_blocks_{i}/downsamplers_0/conv", stride=2, pad=0)
    x = resnet(x, "encoder/mid_block/resnets_0")
    x = attention(x, "encoder/mid_block/attentions_0")
    x = resnet(x, "encoder/mid_block/resnets_1")
    m = conv(conv(F.silu(gn(x, "encoder/conv_norm_out")), "encoder/conv_out"), "quant_conv", pad=0)
    mean, logvar = m.chunk(2, dim=1)
    return torch.cat([mean, logvar.clamp(-30.0, 20.0)], dim=1).permute(0, 2, 3, 1)
