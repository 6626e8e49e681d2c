"""Thin typed wrappers: torch tensors in -> C-ABI calls on the current CUDA stream.

torch is used for device memory and streams only; every arithmetic op below runs in
``libddpo_b200.so``.  Tensors must live on the current CUDA device.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AttentionArgs, DdimCommon, GroupNormArgs, IGemmArgs, check, lib

DDIM_CHUNKS = 8


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    assert t.is_cuda, "ddpo_b200 ops need CUDA tensors (no CPU fallback)"
    return C.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous CUDA {dtype}, got {t.dtype} {t.device}"


# ------------------------------------------------------------------ PRNG --------
def prng_key(seed: int):
    out = (C.c_uint32 * 2)()
    check(lib().ddpo_prng_key_host(C.c_uint64(seed & (2 ** 64 - 1)), out), "prng_key")
    return (int(out[0]), int(out[1]))


def threefry_split(key, num=2):
    k = (C.c_uint32 * 2)(key[0], key[1])
    out = (C.c_uint32 * (2 * num))()
    check(lib().ddpo_threefry_split_host(k, num, out), "threefry_split")
    return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(num)]


def key_tensor(keys, device):
    """[(k0,k1), ...] -> int64-free uint32 storage as int32 tensor [len, 2] on device."""
    flat = []
    for k in keys:
        flat += [k[0] if k[0] < 2 ** 31 else k[0] - 2 ** 32, k[1] if k[1] < 2 ** 31 else k[1] - 2 ** 32]
    return torch.tensor(flat, dtype=torch.int32).view(-1, 2).to(device)


def threefry_normal(key_dev, out):
    _chk(out, torch.float32, "out")
    check(lib().ddpo_threefry_normal(_p(key_dev), _p(out), out.numel(), _stream()), "threefry_normal")
    return out


# ------------------------------------------------------------------ DDIM --------
def ddim_workspace(batch, device):
    return torch.zeros(batch * DDIM_CHUNKS + batch, dtype=torch.float32, device=device)


def _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws):
    b = sample.shape[0]
    n = sample.numel() // b
    for t, nm in ((eps_u, "eps_u"), (eps_c, "eps_c"), (sample, "sample"), (alphas_cumprod, "alphas_cumprod")):
        _chk(t, torch.float32, nm)
    _chk(timesteps, torch.int32, "timesteps")
    assert timesteps.numel() in (1, b)
    c = DdimCommon(_p(eps_u), _p(eps_c), _p(sample), _p(alphas_cumprod), _p(timesteps),
                   1 if timesteps.numel() == b and b > 1 else (0 if timesteps.numel() == 1 else 1),
                   float(final_alpha), int(step_ratio), float(guidance), float(eta), b, n, _p(ws))
    return c


def ddim_step_sample(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta,
                     key_dev, prev_out, logp_out, ws):
    c = _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws)
    check(lib().ddpo_ddim_step_sample(C.byref(c), _p(key_dev), _p(prev_out), _p(logp_out), _stream()), "ddim_step_sample")


def ddim_logprob_fwd(eps_u, eps_c, sample, prev, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta,
                     logp_out, ws):
    c = _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws)
    check(lib().ddpo_ddim_logprob_fwd(C.byref(c), _p(prev), _p(logp_out), _stream()), "ddim_logprob_fwd")


def ddim_logprob_bwd(eps_u, eps_c, sample, prev, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta,
                     dlogp, d_eps_u, d_eps_c, ws):
    c = _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws)
    check(lib().ddpo_ddim_logprob_bwd(C.byref(c), _p(prev), _p(dlogp), _p(d_eps_u), _p(d_eps_c), _stream()),
          "ddim_logprob_bwd")


def ppo_loss(logp, old_logp, adv, clip_range, info_out, dlogp_out):
    check(lib().ddpo_ppo_loss(_p(logp), _p(old_logp), _p(adv), logp.numel(), float(clip_range), _p(info_out),
                              _p(dlogp_out), _stream()), "ppo_loss")


# ------------------------------------------------------------------ GEMM --------
def igemm(*, a0, wt, n, a1=None, c0=None, c1=0, lda0=None, lda1=None, conv=None, m=None, taps=1, stride=1,
          bias=None, rowvec=None, rows_per_sample=0, rowvec_ld=0, residual=None, ld_res=0, out_f32=None,
          out_bf16=None, ld_out=0, geglu=False, accumulate=False, bn=0):
    """conv=(batch, h_out, w_out) for convolutions, else linear with m rows."""
    a = IGemmArgs()
    a.a0, a.a1 = _p(a0), _p(a1)
    a.c0 = int(c0 if c0 is not None else a0.shape[-1])
    a.c1 = int(c1)
    a.lda0 = int(lda0 if lda0 is not None else a.c0)
    a.lda1 = int(lda1 if lda1 is not None else a.c1)
    if conv is not None:
        a.is_conv, a.batch, a.h, a.w = 1, int(conv[0]), int(conv[1]), int(conv[2])
        a.m = 0
    else:
        a.is_conv, a.batch, a.h, a.w = 0, 0, 0, 0
        a.m = int(m)
    a.conv_stride, a.taps, a.n = int(stride), int(taps), int(n)
    a.wt, a.bias, a.rowvec = _p(wt), _p(bias), _p(rowvec)
    a.rows_per_sample, a.rowvec_ld = int(rows_per_sample), int(rowvec_ld)
    a.residual, a.ld_res = _p(residual), int(ld_res)
    a.out_f32, a.out_bf16, a.ld_out = _p(out_f32), _p(out_bf16), int(ld_out)
    a.geglu, a.accumulate_out, a.bn_override = int(geglu), int(accumulate), int(bn)
    check(lib().ddpo_igemm(C.byref(a), _stream()), "igemm")


# ------------------------------------------------------------------ norms -------
def groupnorm_workspace_floats(batch, hw, channels):
    return int(lib().ddpo_groupnorm_workspace_floats(batch, hw, channels))


def _gn_args(x0, x1, c0, c1, batch, hw, scale, bias, silu, y_bf16, y_f32, raw_bf16, ws, ld0=0, ld1=0, skip_stats=False):
    return GroupNormArgs(_p(x0), _p(x1), int(c0), int(c1), int(ld0), int(ld1), int(batch), int(hw), _p(scale),
                         _p(bias), 1e-5, int(silu), _p(y_bf16), _p(y_f32), _p(raw_bf16), _p(ws), int(skip_stats))


def groupnorm_fwd(x0, scale, bias, ws, batch, hw, c0, x1=None, c1=0, silu=True, y_bf16=None, y_f32=None,
                  raw_bf16=None, skip_stats=False):
    a = _gn_args(x0, x1, c0, c1, batch, hw, scale, bias, silu, y_bf16, y_f32, raw_bf16, ws, skip_stats=skip_stats)
    check(lib().ddpo_groupnorm_fwd(C.byref(a), _stream()), "groupnorm_fwd")


def groupnorm_bwd(x0, scale, bias, ws, batch, hw, c0, dy, dx0, dscale, dbias, x1=None, c1=0, dx1=None, silu=True,
                  accumulate=False, ldd0=0, ldd1=0):
    a = _gn_args(x0, x1, c0, c1, batch, hw, scale, bias, silu, None, None, None, ws)
    check(lib().ddpo_groupnorm_bwd(C.byref(a), _p(dy), _p(dx0), _p(dx1), int(ldd0), int(ldd1), int(accumulate),
                                   _p(dscale), _p(dbias), _stream()), "groupnorm_bwd")


def layernorm_fwd(x, scale, bias, y_bf16, m, c, stats=None):
    check(lib().ddpo_layernorm_fwd(_p(x), _p(scale), _p(bias), _p(y_bf16), _p(stats), int(m), int(c), 1e-5, _stream()),
          "layernorm_fwd")


def layernorm_bwd_workspace_floats(m, c):
    return int(lib().ddpo_layernorm_bwd_workspace_floats(m, c))


def layernorm_bwd(x, scale, stats, dy, dx, dscale, dbias, ws, m, c, accumulate=False):
    check(lib().ddpo_layernorm_bwd(_p(x), _p(scale), _p(stats), _p(dy), _p(dx), int(accumulate), _p(dscale),
                                   _p(dbias), _p(ws), int(m), int(c), _stream()), "layernorm_bwd")


# ------------------------------------------------------------ small layers -------
def prep_weight(src, dst, k, n, ldk=None, row_offset=0, col_offset=0, geglu_bn=0):
    check(lib().ddpo_prep_weight(_p(src), _p(dst), int(k), int(n), int(ldk or k), int(row_offset), int(col_offset),
                                 int(geglu_bn), _stream()), "prep_weight")


def prep_weight_dgrad(src, dst, taps, k, n):
    check(lib().ddpo_prep_weight_dgrad(_p(src), _p(dst), int(taps), int(k), int(n), _stream()), "prep_weight_dgrad")


def permute_geglu_bias(src, dst, n, bn):
    check(lib().ddpo_permute_geglu_bias(_p(src), _p(dst), int(n), int(bn), _stream()), "permute_geglu_bias")


def cast_bf16(x, y):
    check(lib().ddpo_cast_bf16(_p(x), _p(y), x.numel(), _stream()), "cast_bf16")


def upsample2x_bf16(x, y, batch, h, w, c):
    check(lib().ddpo_upsample2x_bf16(_p(x), _p(y), batch, h, w, c, _stream()), "upsample2x_bf16")


def upsample2x_bwd(dy, dx, batch, h, w, c, accumulate=False):
    check(lib().ddpo_upsample2x_bwd(_p(dy), _p(dx), batch, h, w, c, int(accumulate), _stream()), "upsample2x_bwd")


def conv_in(x_nchw, w, bias, y_nhwc, batch, cin, h, wd, cout):
    check(lib().ddpo_conv_in(_p(x_nchw), _p(w), _p(bias), _p(y_nhwc), batch, cin, h, wd, cout, _stream()), "conv_in")


def conv_out(x_nhwc, w, bias, y_nchw, batch, h, wd, cin, cout):
    check(lib().ddpo_conv_out(_p(x_nhwc), _p(w), _p(bias), _p(y_nchw), batch, h, wd, cin, cout, _stream()), "conv_out")


def timestep_sincos(t, out, batch, dim):
    check(lib().ddpo_timestep_sincos(_p(t), 1 if t.numel() > 1 else 0, _p(out), batch, dim, _stream()), "timestep_sincos")


def dense_small(x, w, bias, y, batch, k, n, silu_in=False, silu_out=False):
    check(lib().ddpo_dense_small(_p(x), _p(w), _p(bias), _p(y), batch, k, n, int(silu_in), int(silu_out), _stream()),
          "dense_small")


# --------------------------------------------------------------- attention -------
def attention_fwd(q, k, v, out, batch, heads, nq, nk, ldq, ldk, ldv, ldo, lse=None):
    a = AttentionArgs(_p(q), _p(k), _p(v), _p(out), _p(lse), batch, heads, nq, nk, 64, ldq, ldk, ldv, ldo)
    check(lib().ddpo_attention_fwd(C.byref(a), _stream()), "attention_fwd")
