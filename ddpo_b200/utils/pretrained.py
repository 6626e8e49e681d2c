"""Ingest a LOCAL Stable-Diffusion checkpoint directory (the diffusers layout the reference downloads with
``FlaxStableDiffusionPipeline.from_pretrained`` in ``ddpo/utils/serialization.py:336-341``) into this package's flat
Flax-layout parameter buffers:

    <dir>/unet/config.json + diffusion_flax_model.msgpack | diffusion_pytorch_model.safetensors | .bin
    <dir>/vae/config.json  + (same file names)            -- decoder + post_quant_conv are used
    <dir>/text_encoder/config.json + flax_model.msgpack | model.safetensors | pytorch_model.bin
    <dir>/scheduler/scheduler_config.json
    <dir>/tokenizer/                                      -- handed to ``transformers.CLIPTokenizer`` when present

There is no network here, so nothing is downloaded: a hub id resolves only if a directory of that name exists under
``cache`` (``<cache>/<id>`` or ``<cache>/models--org--name/snapshots/*``).  Flax checkpoints are read with the msgpack layout
of flax 0.6.9; PyTorch checkpoints go through the standard name / layout conversion (conv OIHW -> HWIO, Linear
[out, in] -> Dense [in, out], ``weight`` of a norm -> ``scale``, ``to_out.0`` -> ``to_out_0``, ``ff.net.0`` -> ``ff/net_0``).
"""
import glob
import json
import os

import numpy as np
import torch

from .. import unet_spec


def resolve_dir(pretrained_model, cache="cache"):
    """local directory holding the checkpoint, or None"""
    cands = [pretrained_model, os.path.join(cache, pretrained_model),
             os.path.join(cache, "models--" + str(pretrained_model).replace("/", "--"), "snapshots", "*")]
    for c in cands:
        for d in sorted(glob.glob(os.path.expanduser(str(c)))):
            if os.path.isfile(os.path.join(d, "unet", "config.json")):
                return d
    return None


def _read_weights(folder, flax_name, pt_names):
    """-> ("flax", nested dict) or ("pt", flat {name: ndarray})"""
    p = os.path.join(folder, flax_name)
    if os.path.isfile(p):
        from .serialization import msgpack_restore
        with open(p, "rb") as f:
            return "flax", msgpack_restore(f.read())
    for nm in pt_names:
        p = os.path.join(folder, nm)
        if not os.path.isfile(p):
            continue
        if nm.endswith(".safetensors"):
            from safetensors.numpy import load_file
            return "pt", {k: np.asarray(v) for k, v in load_file(p).items()}
        sd = torch.load(p, map_location="cpu", weights_only=True)
        return "pt", {k: v.float().numpy() for k, v in sd.items()}
    raise FileNotFoundError(f"no weights under {folder} (looked for {flax_name}, {', '.join(pt_names)})")


_INDEXED = ("down_blocks", "up_blocks", "resnets", "attentions", "transformer_blocks", "downsamplers", "upsamplers", "net",
            "to_out", "layers")


def pt_key_to_flax(key):
    """'down_blocks.0.attentions.1.transformer_blocks.0.attn1.to_out.0.weight' ->
    ('down_blocks_0/attentions_1/transformer_blocks_0/attn1/to_out_0', 'weight')"""
    parts = key.split(".")
    leaf, parts = parts[-1], parts[:-1]
    out, i = [], 0
    while i < len(parts):
        if parts[i] in _INDEXED and i + 1 < len(parts) and parts[i + 1].isdigit():
            out.append(f"{parts[i]}_{parts[i + 1]}" if parts[i] != "layers" else f"layers/{parts[i + 1]}")
            i += 2
        else:
            out.append(parts[i])
            i += 1
    return "/".join(out), leaf


def pt_state_to_flax_flat(sd, table, embeddings=()):
    """PyTorch state dict -> {flax name: ndarray in Flax layout} for the names of ``table`` (name -> (offset, shape))."""
    out = {}
    for key, v in sd.items():
        base, leaf = pt_key_to_flax(key)
        v = np.asarray(v, np.float32)
        if leaf == "weight":
            for cand, conv in ((base + "/kernel", True), (base + "/scale", False), (base + "/embedding", False)):
                if cand in table:
                    shape = table[cand][1]
                    if not conv:
                        w = v
                    elif v.ndim == 4:
                        w = v.transpose(2, 3, 1, 0)                   # OIHW -> HWIO
                        if len(shape) == 2:                           # a 1x1 conv checkpoint into a Dense layer
                            w = w.reshape(shape)
                    else:
                        w = v.T                                       # [out, in] -> [in, out]
                        if len(shape) == 4:                           # a Dense checkpoint into a 1x1 conv
                            w = w.reshape(shape)
                    out[cand] = w
                    break
        elif leaf == "bias" and base + "/bias" in table:
            out[base + "/bias"] = v
    return out


def _flatten_tree(tree, prefix=""):
    flat = {}
    for k, v in tree.items():
        name = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict):
            flat.update(_flatten_tree(v, name))
        else:
            flat[name] = np.asarray(v, np.float32)
    return flat


def fill_flat(table, total, named, what):
    flat = torch.zeros(total, dtype=torch.float32)
    missing = [n for n in table if n not in named]
    if missing:
        raise KeyError(f"{what}: {len(missing)} parameters missing from the checkpoint, e.g. {missing[:4]}")
    for name, (off, shape) in table.items():
        a = np.asarray(named[name], np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{what}: {name} has shape {a.shape} in the checkpoint, expected {tuple(shape)}")
        flat[off:off + a.size] = torch.from_numpy(np.ascontiguousarray(a).reshape(-1))
    return flat


def unet_config_from_json(cfg_json):
    boc = tuple(cfg_json["block_out_channels"])
    ahd = cfg_json.get("attention_head_dim", 8)
    ahd = tuple(ahd) if isinstance(ahd, (list, tuple)) else (int(ahd),) * len(boc)
    down = cfg_json.get("down_block_types", ("CrossAttnDownBlock2D",) * (len(boc) - 1) + ("DownBlock2D",))
    return unet_spec.UNetConfig(
        in_channels=cfg_json.get("in_channels", 4), out_channels=cfg_json.get("out_channels", 4), block_out_channels=boc,
        layers_per_block=cfg_json.get("layers_per_block", 2), attention_head_dim=ahd,
        cross_attention_dim=cfg_json.get("cross_attention_dim", 1024),
        down_has_attn=tuple("CrossAttn" in t for t in down), use_linear_projection=bool(cfg_json.get("use_linear_projection", False)),
        sample_size=cfg_json.get("sample_size", 64))


def load_unet_weights(directory):
    """-> (UNetConfig, flat fp32 parameters in Flax layout)"""
    folder = os.path.join(directory, "unet")
    with open(os.path.join(folder, "config.json")) as f:
        cfg = unet_config_from_json(json.load(f))
    table, total = unet_spec.param_offsets(cfg)
    kind, w = _read_weights(folder, "diffusion_flax_model.msgpack",
                            ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"))
    named = _flatten_tree(w) if kind == "flax" else pt_state_to_flax_flat(w, table)
    return cfg, fill_flat(table, total, named, "unet")


def load_vae_decoder_weights(directory, part="decoder"):
    """-> (VAEConfig, flat fp32 parameters of the decoder (+ post_quant_conv) or, part="encoder", the encoder (+ quant_conv))"""
    from .. import vae as V
    folder = os.path.join(directory, "vae")
    with open(os.path.join(folder, "config.json")) as f:
        j = json.load(f)
    cfg = V.VAEConfig(latent_channels=j.get("latent_channels", 4), out_channels=j.get("out_channels", 3),
                      block_out_channels=tuple(j["block_out_channels"]), layers_per_block=j.get("layers_per_block", 2),
                      sample_size=j.get("sample_size", 512) // 2 ** (len(j["block_out_channels"]) - 1))
    table, total = V.param_offsets(cfg, part=part)
    kind, w = _read_weights(folder, "diffusion_flax_model.msgpack",
                            ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"))
    if kind == "flax":
        named = _flatten_tree(w)
    else:
        # PyTorch VAE names differ in one place: the mid-block attention (to_q/to_k/to_v/to_out.0 in recent diffusers,
        # query/key/value/proj_attn in 0.12) -- accept both
        ren = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}
        w2 = {}
        for k, v in w.items():
            for a, b in ren.items():
                k = k.replace(f"attentions.0.{a}.", f"attentions.0.{b}.")
            w2[k] = v
        named = pt_state_to_flax_flat(w2, table)
    return cfg, fill_flat(table, total, named, f"vae {part}")


def load_text_encoder_weights(directory):
    from .. import text_encoder as T
    folder = os.path.join(directory, "text_encoder")
    with open(os.path.join(folder, "config.json")) as f:
        j = json.load(f)
    cfg = T.CLIPTextConfig(vocab_size=j["vocab_size"], hidden_size=j["hidden_size"], intermediate_size=j["intermediate_size"],
                           num_hidden_layers=j["num_hidden_layers"], num_attention_heads=j["num_attention_heads"],
                           max_position_embeddings=j.get("max_position_embeddings", 77), hidden_act=j.get("hidden_act", "quick_gelu"))
    table, total = T.param_offsets(cfg)
    kind, w = _read_weights(folder, "flax_model.msgpack", ("model.safetensors", "pytorch_model.bin"))
    named = _flatten_tree(w) if kind == "flax" else pt_state_to_flax_flat(w, table)
    return cfg, fill_flat(table, total, named, "text encoder")


def load_scheduler_config(directory):
    p = os.path.join(directory, "scheduler", "scheduler_config.json")
    if not os.path.isfile(p):
        return {}
    with open(p) as f:
        j = json.load(f)
    keep = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "trained_betas", "set_alpha_to_one",
            "steps_offset", "prediction_type")
    return {k: j[k] for k in keep if k in j}


def load_tokenizer(directory):
    p = os.path.join(directory, "tokenizer")
    if not os.path.isdir(p):
        return None
    from transformers import CLIPTokenizer
    return CLIPTokenizer.from_pretrained(p)
