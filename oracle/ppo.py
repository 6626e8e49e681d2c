"""NumPy restatement of the PPO clipped surrogate in the reference's
``ddpo/training/policy_gradient.py:60,121-134`` (loss, info) with its analytic
gradient w.r.t. ``log_prob``.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

ADV_CLIP_MAX = 10.0
f32 = np.float32


def cfg_combine(uncond, cond, guidance_scale):
    """``training/policy_gradient.py:103-105`` / ``pipeline_flax_stable_diffusion.py:226-229``."""
    return (uncond + f32(guidance_scale) * (cond - uncond)).astype(f32)


def ppo_loss(log_prob, old_log_prob, advantages, clip_range):
    lp = np.asarray(log_prob, f32)
    old = np.asarray(old_log_prob, f32)
    adv = np.clip(np.asarray(advantages, f32), -ADV_CLIP_MAX, ADV_CLIP_MAX).astype(f32)
    ratio = np.exp(lp - old).astype(f32)
    unclipped = -adv * ratio
    clipped = -adv * np.clip(ratio, f32(1.0 - clip_range), f32(1.0 + clip_range))
    loss = np.mean(np.maximum(unclipped, clipped), dtype=f32)
    info = {
        "approx_kl": f32(0.5) * np.mean((lp - old) ** 2, dtype=f32),
        "clipfrac": np.mean((np.abs(ratio - f32(1.0)) > f32(clip_range)).astype(f32), dtype=f32),
        "loss": loss,
    }
    # d loss / d log_prob: max() picks the unclipped branch when unclipped >= clipped
    # (ties -> unclipped, as jnp.maximum's gradient splits evenly only on exact ties of
    #  *different* expressions; on the tie both branches have the same value and, inside
    #  the clip interval, the same derivative).
    use_unclipped = unclipped >= clipped
    inside = (ratio >= f32(1.0 - clip_range)) & (ratio <= f32(1.0 + clip_range))
    d_unclipped = -adv * ratio
    d_clipped = np.where(inside, -adv * ratio, f32(0))
    dlp = np.where(use_unclipped, d_unclipped, d_clipped) / f32(lp.shape[0])
    return loss, info, dlp.astype(f32)


def per_prompt_advantages_global(rewards):
    """``pipeline/policy_gradient.py:347`` (no epsilon on the std)."""
    r = np.asarray(rewards)
    return (r - np.mean(r)) / np.std(r)
