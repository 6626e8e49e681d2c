"""``utils/pretrained.py``: a local diffusers checkpoint directory -> flat Flax-layout buffers.  No real checkpoint exists
offline, so the directories are synthesised: Flax msgpack from this package's own trees, and PyTorch safetensors whose KEYS
come from third-party module trees (``transformers.CLIPTextModel``) or from the PyTorch-idiom twin (diffusers naming)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ddpo_b200 import text_encoder as T, unet_spec, vae as V  # noqa: E402
from ddpo_b200.utils import pretrained as P, serialization as S  # noqa: E402


def _tree(named):
    tree = {}
    for name, v in named.items():
        node = tree
        parts = name.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.asarray(v, np.float32)
    return tree


def _unet_json(cfg):
    return {"block_out_channels": list(cfg.block_out_channels), "attention_head_dim": list(cfg.attention_head_dim),
            "cross_attention_dim": cfg.cross_attention_dim, "use_linear_projection": cfg.use_linear_projection,
            "sample_size": cfg.sample_size, "layers_per_block": cfg.layers_per_block, "in_channels": 4, "out_channels": 4,
            "down_block_types": ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"]}


@pytest.mark.parametrize("fmt", ["flax", "pytorch"])
def test_unet_checkpoint_directory_round_trip(tmp_path, fmt):
    from _torch_twin import UNet2DConditionTwin, load_flax_params
    from safetensors.torch import save_file
    cfg = unet_spec.TINY
    flat = unet_spec.init_flat_params(cfg, 7)
    views = unet_spec.views(flat, cfg)
    d = tmp_path / "ckpt"
    os.makedirs(d / "unet")
    os.makedirs(d / "scheduler")
    with open(d / "unet" / "config.json", "w") as f:
        json.dump(_unet_json(cfg), f)
    with open(d / "scheduler" / "scheduler_config.json", "w") as f:
        json.dump({"num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012, "beta_schedule": "scaled_linear",
                   "set_alpha_to_one": False, "steps_offset": 1, "prediction_type": "v_prediction", "clip_sample": False}, f)
    if fmt == "flax":
        with open(d / "unet" / "diffusion_flax_model.msgpack", "wb") as f:
            f.write(S.msgpack_serialize(_tree({k: v.numpy() for k, v in views.items()})))
    else:
        twin = load_flax_params(UNet2DConditionTwin(cfg), views)          # keys = the PyTorch model's own names
        save_file({k: v.contiguous() for k, v in twin.state_dict().items()}, str(d / "unet" / "diffusion_pytorch_model.safetensors"))
    assert P.resolve_dir(str(d)) == str(d)
    cfg2, flat2 = P.load_unet_weights(str(d))
    assert cfg2 == cfg
    assert torch.equal(flat2, flat)
    assert P.load_scheduler_config(str(d))["prediction_type"] == "v_prediction"


def test_text_encoder_checkpoint_from_transformers_state_dict(tmp_path):
    """keys written by transformers' own CLIPTextModel -> our Flax-layout buffer -> same hidden states as that model"""
    tr = pytest.importorskip("transformers")
    from safetensors.torch import save_file
    from oracle import text_encoder as OT
    cfg = T.TEXT_TINY
    hc = tr.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           max_position_embeddings=77, hidden_act=cfg.hidden_act, layer_norm_eps=cfg.layer_norm_eps,
                           eos_token_id=2, bos_token_id=0, pad_token_id=1)
    torch.manual_seed(0)
    model = tr.CLIPTextModel(hc).eval()
    d = tmp_path / "ckpt"
    os.makedirs(d / "text_encoder")
    os.makedirs(d / "unet")
    with open(d / "unet" / "config.json", "w") as f:
        json.dump(_unet_json(unet_spec.TINY), f)
    with open(d / "text_encoder" / "config.json", "w") as f:
        json.dump({"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
                   "num_hidden_layers": cfg.num_hidden_layers, "num_attention_heads": cfg.num_attention_heads,
                   "max_position_embeddings": 77, "hidden_act": cfg.hidden_act}, f)
    save_file({k: v.contiguous() for k, v in model.state_dict().items() if v.dtype.is_floating_point},
              str(d / "text_encoder" / "model.safetensors"))
    cfg2, flat = P.load_text_encoder_weights(str(d))
    assert cfg2.hidden_size == cfg.hidden_size and cfg2.num_hidden_layers == cfg.num_hidden_layers
    ids = torch.randint(3, cfg.vocab_size, (2, 77), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = model(input_ids=ids).last_hidden_state
    got = OT.encode(T.views(flat, cfg2), cfg2, ids)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-4, atol=2e-5)


def test_vae_decoder_checkpoint_round_trip(tmp_path):
    cfg = V.VAE_MICRO
    flat = V.init_flat_params(cfg, 2)
    views = V.views(flat, cfg)
    d = tmp_path / "ckpt"
    os.makedirs(d / "vae")
    with open(d / "vae" / "config.json", "w") as f:
        json.dump({"block_out_channels": list(cfg.block_out_channels), "layers_per_block": cfg.layers_per_block,
                   "latent_channels": 4, "out_channels": 3, "sample_size": cfg.sample_size * 8}, f)
    tree = _tree({k: v.numpy() for k, v in views.items()})
    tree["encoder"] = {"conv_in": {"kernel": np.zeros((3, 3, 3, 8), np.float32)}}      # ignored half of the checkpoint
    with open(d / "vae" / "diffusion_flax_model.msgpack", "wb") as f:
        f.write(S.msgpack_serialize(tree))
    cfg2, flat2 = P.load_vae_decoder_weights(str(d))
    assert cfg2.block_out_channels == cfg.block_out_channels and cfg2.sample_size == cfg.sample_size
    assert torch.equal(flat2, flat)


def test_random_init_is_an_explicit_opt_in(monkeypatch, tmp_path):
    monkeypatch.delenv(S.ALLOW_RANDOM_ENV, raising=False)
    with pytest.raises(FileNotFoundError):
        S.load_unet(None, pretrained_model="stabilityai/stable-diffusion-2-base", cache=str(tmp_path), device="cpu")
    with pytest.raises(KeyError):
        P.fill_flat({"a/kernel": (0, (2, 2))}, 4, {}, "unet")
