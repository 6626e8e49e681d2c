"""CPU tests of the text-encoder spec and oracle.  The oracle restatement is PINNED against the installed
``transformers`` library's own ``CLIPTextModel`` (PyTorch implementation of the architecture the reference loads as
``FlaxCLIPTextModel``) with identical random weights."""
import numpy as np
import pytest
import torch

from ddpo_b200 import text_encoder as T
from oracle import text_encoder as OT


def test_parameter_counts_match_published_towers():
    # CLIP ViT-L/14 text tower (SD1.x): 123 060 480 parameters; OpenCLIP ViT-H text tower with 23 layers (SD2): 340 387 840
    assert T.num_params(T.SD1_TEXT) == 123_060_480
    assert T.num_params(T.SD2_TEXT) == 340_387_840


def _hf_model(cfg, flat):
    tr = pytest.importorskip("transformers")
    hc = tr.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           max_position_embeddings=cfg.max_position_embeddings, hidden_act=cfg.hidden_act,
                           layer_norm_eps=cfg.layer_norm_eps, attention_dropout=0.0, eos_token_id=2, bos_token_id=0,
                           pad_token_id=1)
    model = tr.CLIPTextModel(hc).eval()
    p = T.views(flat, cfg)
    sd = model.state_dict()
    put = lambda k, v: sd[k].copy_(v)
    with torch.no_grad():
        put("text_model.embeddings.token_embedding.weight", p["text_model/embeddings/token_embedding/embedding"])
        put("text_model.embeddings.position_embedding.weight", p["text_model/embeddings/position_embedding/embedding"])
        for i in range(cfg.num_hidden_layers):
            b, hb = f"text_model/encoder/layers/{i}", f"text_model.encoder.layers.{i}"
            for ln in ("layer_norm1", "layer_norm2"):
                put(f"{hb}.{ln}.weight", p[f"{b}/{ln}/scale"])
                put(f"{hb}.{ln}.bias", p[f"{b}/{ln}/bias"])
            for m_ in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "mlp.fc1", "mlp.fc2"):
                fm = m_.replace(".", "/")
                put(f"{hb}.{m_}.weight", p[f"{b}/{fm}/kernel"].t())      # torch Linear stores [out, in]
                put(f"{hb}.{m_}.bias", p[f"{b}/{fm}/bias"])
        put("text_model.final_layer_norm.weight", p["text_model/final_layer_norm/scale"])
        put("text_model.final_layer_norm.bias", p["text_model/final_layer_norm/bias"])
    model.load_state_dict(sd)
    return model


@pytest.mark.parametrize("cfg", [T.TEXT_TINY, T.TEXT_TINY_QUICK])
def test_oracle_matches_transformers_clip_text_model(cfg):
    flat = T.init_flat_params(cfg, 0)
    model = _hf_model(cfg, flat)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(3, cfg.vocab_size, (3, 77), generator=g)
    with torch.no_grad():
        ref = model(input_ids=ids).last_hidden_state
    got = OT.encode(T.views(flat, cfg), cfg, ids)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-4, atol=2e-5)


def test_oracle_is_causal():
    """changing a later token must not change earlier positions"""
    cfg = T.TEXT_TINY
    flat = T.init_flat_params(cfg, 2)
    ids = torch.randint(3, cfg.vocab_size, (1, 77), generator=torch.Generator().manual_seed(3))
    a = OT.encode(T.views(flat, cfg), cfg, ids)
    ids2 = ids.clone()
    ids2[0, 40] = (ids2[0, 40] + 1) % cfg.vocab_size
    b = OT.encode(T.views(flat, cfg), cfg, ids2)
    assert torch.equal(a[0, :40], b[0, :40]) and not torch.allclose(a[0, 40:], b[0, 40:])


def test_text_encoder_host_assembly_dry_run_against_oracle(monkeypatch):
    """ddpo_b200/text_encoder.py's kernel sequencing on the torch-CPU emulation of the ops it calls (fused q/k/v GEMM,
    causal attention over column-offset views, bias / residual epilogues, activation, final fp32 LayerNorm)."""
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    import _cpu_ops_emulator as E
    monkeypatch.setattr(T, "ops", E)
    monkeypatch.setattr(T, "Arena", E.CpuArena)
    for cfg in (T.TEXT_TINY, T.TEXT_TINY_QUICK):
        flat = T.init_flat_params(cfg, 0)
        enc = T.CLIPTextEncoder(cfg, flat, device="cpu")
        ids = torch.randint(3, cfg.vocab_size, (2, 77), generator=torch.Generator().manual_seed(5))
        got = enc(ids.numpy())[0]
        ref = OT.encode(T.views(flat, cfg), cfg, ids)
        assert got.shape == (2, 77, cfg.hidden_size) and got.dtype == torch.float32
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 2e-2, rel
