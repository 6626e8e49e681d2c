"""CPU tests of the aesthetic reward model: spec, oracle (PINNED against transformers' CLIPVisionModelWithProjection),
preprocessing, and the host assembly of ddpo_b200/clip_vision.py dry-run on the CPU ops emulator."""
import numpy as np
import pytest
import torch

from ddpo_b200 import clip_vision as CV
from oracle import clip_vision as OCV


def test_vit_l14_parameter_count():
    # CLIP ViT-L/14: vision_model 303 179 776 parameters + visual_projection 1024 x 768 = 786 432
    assert CV.num_params(CV.VIT_L14) == 303_179_776 + 786_432
    head = CV.num_params(CV.VIT_L14, head=True) - CV.num_params(CV.VIT_L14)
    assert head == 768 * 1024 + 1024 + 1024 * 128 + 128 + 128 * 64 + 64 + 64 * 16 + 16 + 16 + 1


def _hf_model(cfg, flat):
    tr = pytest.importorskip("transformers")
    hc = tr.CLIPVisionConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                             num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                             image_size=cfg.image_size, patch_size=cfg.patch_size, projection_dim=cfg.projection_dim,
                             hidden_act=cfg.hidden_act, layer_norm_eps=cfg.layer_norm_eps, attention_dropout=0.0)
    model = tr.CLIPVisionModelWithProjection(hc).eval()
    p = CV.views(flat, cfg)
    sd = model.state_dict()
    with torch.no_grad():
        sd["vision_model.embeddings.class_embedding"].copy_(p["vision_model/embeddings/class_embedding"])
        sd["vision_model.embeddings.patch_embedding.weight"].copy_(
            p["vision_model/embeddings/patch_embedding/kernel"].permute(3, 2, 0, 1))                  # HWIO -> OIHW
        sd["vision_model.embeddings.position_embedding.weight"].copy_(p["vision_model/embeddings/position_embedding/embedding"])
        for hf, fx in (("pre_layrnorm", "pre_layrnorm"), ("post_layernorm", "post_layernorm")):
            sd[f"vision_model.{hf}.weight"].copy_(p[f"vision_model/{fx}/scale"])
            sd[f"vision_model.{hf}.bias"].copy_(p[f"vision_model/{fx}/bias"])
        for i in range(cfg.num_hidden_layers):
            b, hb = f"vision_model/encoder/layers/{i}", f"vision_model.encoder.layers.{i}"
            for ln in ("layer_norm1", "layer_norm2"):
                sd[f"{hb}.{ln}.weight"].copy_(p[f"{b}/{ln}/scale"])
                sd[f"{hb}.{ln}.bias"].copy_(p[f"{b}/{ln}/bias"])
            for m_ in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "mlp.fc1", "mlp.fc2"):
                fm = m_.replace(".", "/")
                sd[f"{hb}.{m_}.weight"].copy_(p[f"{b}/{fm}/kernel"].t())
                sd[f"{hb}.{m_}.bias"].copy_(p[f"{b}/{fm}/bias"])
        sd["visual_projection.weight"].copy_(p["visual_projection/kernel"].t())
    model.load_state_dict(sd)
    return model


def test_oracle_matches_transformers_clip_vision_model():
    cfg = CV.VIT_TINY
    flat = CV.init_flat_params(cfg, 0)
    model = _hf_model(cfg, flat)
    x = torch.randn(3, cfg.image_size, cfg.image_size, 3, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = model(pixel_values=x.permute(0, 3, 1, 2)).image_embeds
    got = OCV.image_features(CV.views(flat, cfg), cfg, x)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-4, atol=2e-5)


def test_preprocess_matches_clip_processor_convention():
    img = np.random.default_rng(0).random((2, 64, 64, 3)).astype(np.float32)
    px = CV.preprocess(img, 32)
    assert px.shape == (2, 32, 32, 3) and px.dtype == np.float32
    same = CV.preprocess(img, 64)            # no resize: exactly (uint8 / 255 - mean) / std
    want = ((img * 255).astype(np.uint8).astype(np.float32) / 255 - CV.CLIP_MEAN) / CV.CLIP_STD
    np.testing.assert_allclose(same, want, rtol=1e-6, atol=1e-6)
    tr = pytest.importorskip("transformers")
    proc = tr.CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})
    ref = proc(images=[(i * 255).astype(np.uint8) for i in img], return_tensors="np")["pixel_values"]
    np.testing.assert_allclose(px, np.transpose(ref, (0, 2, 3, 1)), rtol=1e-4, atol=2e-2)   # PIL bicubic on both sides


def test_scorer_host_assembly_dry_run_against_oracle(monkeypatch):
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    import _cpu_ops_emulator as E
    monkeypatch.setattr(CV, "ops", E)
    monkeypatch.setattr(CV, "Arena", E.CpuArena)
    cfg = CV.VIT_TINY
    flat = CV.init_flat_params(cfg, 0)
    sc = CV.AestheticScorer(cfg, flat, device="cpu")
    x = torch.randn(3, cfg.image_size, cfg.image_size, 3, generator=torch.Generator().manual_seed(2))
    feats = sc.image_features(x)
    ref = OCV.image_features(CV.views(flat, cfg), cfg, x)
    rel = ((feats - ref).norm() / ref.norm()).item()
    assert feats.shape == (3, cfg.projection_dim) and rel < 2e-2, rel
    # the head: L2 normalise + five Dense layers; VIT_TINY's projection is 64 wide, so give the head matching weights
    p = CV.views(flat, cfg)
    assert p["aesthetic/Dense_0/kernel"].shape == (cfg.projection_dim, 1024)
    s = sc.score_features(feats)
    want = OCV.aesthetic_score(p, ref)
    assert s.shape == (3, 1)
    np.testing.assert_allclose(s.numpy(), want.numpy(), rtol=5e-2, atol=5e-3)
    scores = sc(np.random.default_rng(1).random((5, 80, 80, 3)).astype(np.float32), chunk=2)
    assert scores.shape == (5, 1) and np.isfinite(scores).all()
