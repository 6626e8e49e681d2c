"""torch-CPU restatement of the aesthetic reward model (reference ``ddpo/training/callbacks.py:60-95``,
``ddpo/models/laion.py:7-18``; 3P transformers==4.28.1 ``FlaxCLIPModel.get_image_features``): patch embedding,
class + position embeddings, pre-LN, pre-LN transformer layers (``quick_gelu``), post-LN of the CLS token, visual
projection, L2 normalisation, five activation-free Dense layers.  Flax parameter names.  TEST INFRASTRUCTURE ONLY.

PINNED: ``tests/test_clip_vision_cpu.py`` checks ``image_features`` against the installed ``transformers``
``CLIPVisionModelWithProjection`` loaded with the same random weights."""
import torch

from .text_encoder import _act
from .unet import layer_norm


def image_features(params, cfg, pixel_values_nhwc, dtype=torch.float32):
    p = {k: v.to(dtype) for k, v in params.items()}
    x = torch.as_tensor(pixel_values_nhwc).to(dtype)
    B, S, P = x.shape[0], cfg.image_size, cfg.patch_size
    n = S // P
    D, H = cfg.hidden_size, cfg.num_attention_heads
    d = D // H
    patches = x.reshape(B, n, P, n, P, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, n * n, P * P * 3)
    pe = patches @ p["vision_model/embeddings/patch_embedding/kernel"].reshape(P * P * 3, D)
    cls = p["vision_model/embeddings/class_embedding"].expand(B, 1, D)
    x = torch.cat([cls, pe], dim=1) + p["vision_model/embeddings/position_embedding/embedding"][None]
    L = x.shape[1]
    x = layer_norm(x, p["vision_model/pre_layrnorm/scale"], p["vision_model/pre_layrnorm/bias"], eps=cfg.layer_norm_eps)
    for i in range(cfg.num_hidden_layers):
        b = f"vision_model/encoder/layers/{i}"
        lin = lambda t, nm: t @ p[f"{b}/{nm}/kernel"] + p[f"{b}/{nm}/bias"]
        h = layer_norm(x, p[b + "/layer_norm1/scale"], p[b + "/layer_norm1/bias"], eps=cfg.layer_norm_eps)
        split = lambda t: t.reshape(B, L, H, d).permute(0, 2, 1, 3)
        q, k, v = split(lin(h, "self_attn/q_proj")) * d ** -0.5, split(lin(h, "self_attn/k_proj")), split(lin(h, "self_attn/v_proj"))
        o = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3).reshape(B, L, D)
        x = x + lin(o, "self_attn/out_proj")
        h = layer_norm(x, p[b + "/layer_norm2/scale"], p[b + "/layer_norm2/bias"], eps=cfg.layer_norm_eps)
        x = x + lin(_act(lin(h, "mlp/fc1"), cfg.hidden_act), "mlp/fc2")
    pooled = layer_norm(x[:, 0], p["vision_model/post_layernorm/scale"], p["vision_model/post_layernorm/bias"],
                        eps=cfg.layer_norm_eps)
    return pooled @ p["visual_projection/kernel"]


def aesthetic_score(params, feats, n_layers=5):
    p = params
    x = feats / feats.norm(dim=-1, keepdim=True)
    for i in range(n_layers):
        x = x @ p[f"aesthetic/Dense_{i}/kernel"].to(x.dtype) + p[f"aesthetic/Dense_{i}/bias"].to(x.dtype)
    return x
