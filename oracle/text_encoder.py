"""torch-CPU restatement of the CLIP text encoder the reference uses to embed prompts (3P transformers==4.28.1
``FlaxCLIPTextModel``; call sites ``pipeline/policy_gradient.py:185-187``, ``ddpo/training/diffusion.py:45-51``):
embeddings -> pre-LN layers with causal self-attention (q scaled by d^-1/2) and ``fc2(act(fc1(x)))`` -> final LN.
Parameters use the Flax names / ``[in, out]`` kernels.  TEST INFRASTRUCTURE ONLY.

PINNED: ``tests/test_text_encoder_cpu.py`` loads the same random weights into the installed ``transformers``
``CLIPTextModel`` (the PyTorch implementation of the same architecture) and checks this restatement against it."""
import math

import torch

from .unet import layer_norm


def _act(x, name):
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    return torch.nn.functional.gelu(x)          # exact (erf) form


def encode(params, cfg, input_ids, dtype=torch.float32):
    p = {k: v.to(dtype) for k, v in params.items()}
    ids = torch.as_tensor(input_ids).long()
    B, L = ids.shape
    D, H = cfg.hidden_size, cfg.num_attention_heads
    d = D // H
    x = p["text_model/embeddings/token_embedding/embedding"][ids] + \
        p["text_model/embeddings/position_embedding/embedding"][:L][None]
    causal = torch.full((L, L), float("-inf"), dtype=dtype).triu(1)
    for i in range(cfg.num_hidden_layers):
        b = f"text_model/encoder/layers/{i}"
        lin = lambda t, n: t @ p[f"{b}/{n}/kernel"] + p[f"{b}/{n}/bias"]
        h = layer_norm(x, p[b + "/layer_norm1/scale"], p[b + "/layer_norm1/bias"], eps=cfg.layer_norm_eps)
        split = lambda t: t.reshape(B, L, H, d).permute(0, 2, 1, 3)
        q, k, v = split(lin(h, "self_attn/q_proj")) * d ** -0.5, split(lin(h, "self_attn/k_proj")), split(lin(h, "self_attn/v_proj"))
        w = torch.softmax(q @ k.transpose(-1, -2) + causal, dim=-1)
        o = (w @ v).permute(0, 2, 1, 3).reshape(B, L, D)
        x = x + lin(o, "self_attn/out_proj")
        h = layer_norm(x, p[b + "/layer_norm2/scale"], p[b + "/layer_norm2/bias"], eps=cfg.layer_norm_eps)
        x = x + lin(_act(lin(h, "mlp/fc1"), cfg.hidden_act), "mlp/fc2")
    return layer_norm(x, p["text_model/final_layer_norm/scale"], p["text_model/final_layer_norm/bias"],
                      eps=cfg.layer_norm_eps)
