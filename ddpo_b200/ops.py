"""Thin typed wrappers: torch tensors in -> C-ABI calls on the current CUDA stream.

torch is used for device memory and streams only; every arithmetic op below runs in
``libddpo_b200.so``.  Tensors must live on the current CUDA device.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AttentionArgs, DdimCommon, GroupNormArgs, IGemmArgs, check, lib

DDIM_CHUNKS = 8

# ---- instrumentation (bench.py): kernel-launch counter and optional CUDA-event profile ----
LAUNCH_COUNT = 0           # kernels launched through this module since import
PROFILE = None             # list of (name, work, event0, event1) while profiling, else None
_KERNELS_PER_CALL = {"groupnorm_fwd": 2, "groupnorm_bwd": 3, "layernorm_bwd": 2}   # groupnorm_fwd: finalize + apply


PROFILE_TAGS = []          # parallel to PROFILE: a shape tag per entry (bench.py --shapes)


def _run(name, status, work=0.0, ev=None, tag=""):
    global LAUNCH_COUNT
    LAUNCH_COUNT += _KERNELS_PER_CALL.get(name, 1)
    if ev is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((name, work, ev, e1))
        PROFILE_TAGS.append(tag)
    check(status, name)


def _ev():
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


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


def threefry_randint(key, n, minval, maxval):
    """jax.random.randint(key, (n,), minval, maxval) on the host -> list of ints."""
    k = (C.c_uint32 * 2)(key[0], key[1])
    out = (C.c_int32 * int(n))()
    check(lib().ddpo_threefry_randint_host(k, int(n), int(minval), int(maxval), out), "threefry_randint")
    return [int(v) for v in out]


def key_tensor(keys, device):
    """[(k0,k1), ...] -> int64-free uint32 storage as int32 tensor [len, 2] on device."""
    flat = []
    for k in keys:
        flat += [k[0] if k[0] < 2 ** 31 else k[0] - 2 ** 32, k[1] if k[1] < 2 ** 31 else k[1] - 2 ** 32]
    return torch.tensor(flat, dtype=torch.int32).view(-1, 2).to(device)


def threefry_normal(key_dev, out):
    _chk(out, torch.float32, "out")
    _e = _ev()
    _run("threefry_normal", lib().ddpo_threefry_normal(_p(key_dev), _p(out), out.numel(), _stream()), 0.0, _e)
    return out


# ------------------------------------------------------------------ DDIM --------
def ddim_workspace(batch, device):
    return torch.zeros(batch * DDIM_CHUNKS + batch, dtype=torch.float32, device=device)


PREDICTION_TYPES = {"epsilon": 0, "sample": 1, "v_prediction": 2}   # scheduling_ddim_flax.py:303-321


def prediction_code(prediction_type):
    """FlaxDDIMScheduler's config string -> the C ABI's DDPO_PRED_* code; same ValueError text as the reference (:317-321)."""
    if prediction_type not in PREDICTION_TYPES:
        raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or"
                         " `v_prediction`")
    return PREDICTION_TYPES[prediction_type]


def _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws,
                 pred="epsilon"):
    b = sample.shape[0]
    n = sample.numel() // b
    for t, nm in ((eps_u, "eps_u"), (eps_c, "eps_c"), (sample, "sample"), (alphas_cumprod, "alphas_cumprod")):
        _chk(t, torch.float32, nm)
    _chk(timesteps, torch.int32, "timesteps")
    assert timesteps.numel() in (1, b)
    c = DdimCommon(_p(eps_u), _p(eps_c), _p(sample), _p(alphas_cumprod), _p(timesteps),
                   1 if timesteps.numel() == b and b > 1 else (0 if timesteps.numel() == 1 else 1),
                   float(final_alpha), int(step_ratio), float(guidance), float(eta), b, n, _p(ws), prediction_code(pred))
    return c


def ddim_step_sample(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta,
                     key_dev, prev_out, logp_out, ws, pred="epsilon"):
    c = _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws, pred)
    _e = _ev()
    _run("ddim_step_sample", lib().ddpo_ddim_step_sample(C.byref(c), _p(key_dev), _p(prev_out), _p(logp_out), _stream()), 0.0, _e)


def ddim_logprob_fwd(eps_u, eps_c, sample, prev, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta,
                     logp_out, ws, pred="epsilon"):
    c = _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws, pred)
    _e = _ev()
    _run("ddim_logprob_fwd", lib().ddpo_ddim_logprob_fwd(C.byref(c), _p(prev), _p(logp_out), _stream()), 0.0, _e)


def ddim_logprob_bwd(eps_u, eps_c, sample, prev, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta,
                     dlogp, d_eps_u, d_eps_c, ws, pred="epsilon"):
    c = _ddim_common(eps_u, eps_c, sample, alphas_cumprod, timesteps, final_alpha, step_ratio, guidance, eta, ws, pred)
    _e = _ev()
    _run("ddim_logprob_bwd", lib().ddpo_ddim_logprob_bwd(C.byref(c), _p(prev), _p(dlogp), _p(d_eps_u), _p(d_eps_c), _stream()), 0.0, _e)


def ppo_loss(logp, old_logp, adv, clip_range, info_out, dlogp_out, micro_batch=None):
    _e = _ev()
    _run("ppo_loss", lib().ddpo_ppo_loss(_p(logp), _p(old_logp), _p(adv), logp.numel(), int(micro_batch or logp.numel()),
                                         float(clip_range), _p(info_out), _p(dlogp_out), _stream()), 0.0, _e)


# ------------------------------------------------------------------ GEMM --------
def igemm(*, a0, wt, n, a1=None, c0=None, c1=0, lda0=None, lda1=None, conv=None, m=None, taps=1, stride=1,
          bias=None, rowvec=None, rows_per_sample=0, rowvec_ld=0, residual=None, ld_res=0, out_f32=None,
          out_bf16=None, ld_out=0, geglu=False, accumulate=False, bn=0, aux_bf16=None, mt=0, pair=0, epi=0, gn_stats=None,
          no_low_pad=False):
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
    a.aux_bf16 = _p(aux_bf16)
    a.mt_override = int(mt)
    a.pair_override = int(pair)
    a.epi_override = int(epi)
    a.gn_stats = _p(gn_stats)
    a.conv_pad = int(no_low_pad)
    rows = a.batch * a.h * a.w if a.is_conv else a.m
    _e = _ev()
    _run("igemm", lib().ddpo_igemm(C.byref(a), _stream()), 2.0 * rows * a.n * a.taps * (a.c0 + a.c1), _e,
         f"M{rows} N{a.n} K{a.taps * (a.c0 + a.c1)} taps{a.taps} s{a.conv_stride}"
         f"{' geglu' if geglu else ''}{' res' if residual is not None else ''}{' f32' if out_f32 is not None else ''}"
         f"{' bf16' if out_bf16 is not None else ''}{' acc' if accumulate else ''}" if _e is not None else "")


# ------------------------------------------------------------------ norms -------
def groupnorm_workspace_floats(batch, hw, channels):
    return int(lib().ddpo_groupnorm_workspace_floats(batch, hw, channels))


def _gn_args(x0, x1, c0, c1, batch, hw, scale, bias, silu, y_bf16, y_f32, raw_bf16, ws, ld0=0, ld1=0, skip_stats=False,
             eps=1e-5, stats0=None, stats1=None, dy_bf16=False):
    return GroupNormArgs(_p(x0), _p(x1), int(c0), int(c1), int(ld0), int(ld1), int(batch), int(hw), _p(scale),
                         _p(bias), float(eps), int(silu), _p(y_bf16), _p(y_f32), _p(raw_bf16), _p(ws), int(skip_stats),
                         _p(stats0), _p(stats1), int(dy_bf16))


def gn_stats_shape(rows, channels):
    """Shape of the slab-statistics buffer an igemm with ``gn_stats=`` fills for its [rows, channels] fp32 output."""
    return ((int(rows) + 31) // 32, int(channels), 2)


def groupnorm_fwd(x0, scale, bias, ws, batch, hw, c0, x1=None, c1=0, silu=True, y_bf16=None, y_f32=None,
                  raw_bf16=None, skip_stats=False, eps=1e-5, stats0=None, stats1=None):
    """``stats0`` / ``stats1``: slab statistics of x0 / x1 left by the igemm that produced them (``gn_stats=``); with both
    (or ``stats0`` alone for one source) and hw % 32 == 0 the statistics pass over x is skipped."""
    a = _gn_args(x0, x1, c0, c1, batch, hw, scale, bias, silu, y_bf16, y_f32, raw_bf16, ws, skip_stats=skip_stats,
                 eps=eps, stats0=stats0, stats1=stats1)
    _e = _ev()
    per = 4 + (2 if y_bf16 is not None else 0) + (4 if y_f32 is not None else 0) + (2 if raw_bf16 is not None else 0)
    _run("groupnorm_fwd", lib().ddpo_groupnorm_fwd(C.byref(a), _stream()), float(batch) * hw * (c0 + c1) * per, _e)


def groupnorm_bwd(x0, scale, bias, ws, batch, hw, c0, dy, dx0, dscale, dbias, x1=None, c1=0, dx1=None, silu=True,
                  accumulate=False, ldd0=0, ldd1=0):
    """``dy`` fp32 or bf16 (the dtype selects the kernel's load path): the dgrad GEMM in front of a GroupNorm writes bf16."""
    assert dy.dtype in (torch.float32, torch.bfloat16)
    a = _gn_args(x0, x1, c0, c1, batch, hw, scale, bias, silu, None, None, None, ws, dy_bf16=dy.dtype == torch.bfloat16)
    _e = _ev()
    _run("groupnorm_bwd", lib().ddpo_groupnorm_bwd(C.byref(a), _p(dy), _p(dx0), _p(dx1), int(ldd0), int(ldd1), int(accumulate),
                                   _p(dscale), _p(dbias), _stream()), 0.0, _e)


def layernorm_fwd(x, scale, bias, y_bf16, m, c, stats=None):
    _e = _ev()
    _run("layernorm_fwd", lib().ddpo_layernorm_fwd(_p(x), _p(scale), _p(bias), _p(y_bf16), _p(stats), int(m), int(c), 1e-5, _stream()), 0.0, _e)


def layernorm_bwd_workspace_floats(m, c):
    return int(lib().ddpo_layernorm_bwd_workspace_floats(m, c))


def layernorm_bwd(x, scale, stats, dy, dx, dscale, dbias, ws, m, c, accumulate=False):
    assert dy.dtype in (torch.float32, torch.bfloat16)
    _e = _ev()
    _run("layernorm_bwd", lib().ddpo_layernorm_bwd(_p(x), _p(scale), _p(stats), _p(dy), _p(dx), int(accumulate), _p(dscale),
                                   _p(dbias), _p(ws), int(m), int(c), int(dy.dtype == torch.bfloat16), _stream()), 0.0, _e)


# ------------------------------------------------------------ small layers -------
def prep_weight(src, dst, k, n, ldk=None, row_offset=0, col_offset=0, geglu_bn=0):
    _e = _ev()
    _run("prep_weight", lib().ddpo_prep_weight(_p(src), _p(dst), int(k), int(n), int(ldk or k), int(row_offset), int(col_offset),
                                 int(geglu_bn), _stream()), 0.0, _e)


def prep_weight_dgrad(src, dst, taps, k, n, ld_dst=0, col_offset=0):
    _e = _ev()
    _run("prep_weight_dgrad", lib().ddpo_prep_weight_dgrad(_p(src), _p(dst), int(taps), int(k), int(n), int(ld_dst),
                                                           int(col_offset), _stream()), 0.0, _e)


def permute_geglu_bias(src, dst, n, bn):
    _e = _ev()
    _run("permute_geglu_bias", lib().ddpo_permute_geglu_bias(_p(src), _p(dst), int(n), int(bn), _stream()), 0.0, _e)


def cast_bf16(x, y):
    _e = _ev()
    _run("cast_bf16", lib().ddpo_cast_bf16(_p(x), _p(y), x.numel(), _stream()), 0.0, _e)


def upsample2x_bf16(x, y, batch, h, w, c):
    _e = _ev()
    _run("upsample2x_bf16", lib().ddpo_upsample2x_bf16(_p(x), _p(y), batch, h, w, c, _stream()), 0.0, _e)


def upsample2x_bwd(dy, dx, batch, h, w, c, accumulate=False):
    _e = _ev()
    _run("upsample2x_bwd", lib().ddpo_upsample2x_bwd(_p(dy), _p(dx), batch, h, w, c, int(accumulate), _stream()), 0.0, _e)


def conv_in(x_nchw, w, bias, y_nhwc, batch, cin, h, wd, cout, gn_stats=None):
    _e = _ev()
    _run("conv_in", lib().ddpo_conv_in(_p(x_nchw), _p(w), _p(bias), _p(y_nhwc), batch, cin, h, wd, cout, _p(gn_stats),
                                       _stream()), 0.0, _e)


def conv_out(x_nhwc, w, bias, y_nchw, batch, h, wd, cin, cout):
    _e = _ev()
    _run("conv_out", lib().ddpo_conv_out(_p(x_nhwc), _p(w), _p(bias), _p(y_nchw), batch, h, wd, cin, cout, _stream()), 0.0, _e)


def timestep_sincos(t, out, batch, dim):
    _e = _ev()
    _run("timestep_sincos", lib().ddpo_timestep_sincos(_p(t), 1 if t.numel() > 1 else 0, _p(out), batch, dim, _stream()), 0.0, _e)


def dense_small(x, w, bias, y, batch, k, n, silu_in=False, silu_out=False):
    _e = _ev()
    _run("dense_small", lib().ddpo_dense_small(_p(x), _p(w), _p(bias), _p(y), batch, k, n, int(silu_in), int(silu_out), _stream()), 0.0, _e)


def dense_small_group_table(entries, device):
    """entries: [(w_off, bias_off, y_off, n)] in floats -> (device uint8 table, total CTAs) for dense_small_grouped."""
    import numpy as np
    rec = np.zeros(len(entries), dtype=[("w", "<i8"), ("b", "<i8"), ("y", "<i8"), ("n", "<i4"), ("cta0", "<i4")])
    cta = 0
    for i, (w_off, b_off, y_off, n) in enumerate(entries):
        rec[i] = (w_off, b_off, y_off, n, cta)
        cta += (n + 31) // 32
    return torch.from_numpy(rec.view(np.uint8).copy()).to(device), cta


def dense_small_grouped(x, params_base, y_base, table, n_groups, total_ctas, batch, k):
    _e = _ev()
    _run("dense_small", lib().ddpo_dense_small_grouped(_p(x), _p(params_base), _p(y_base), _p(table), int(n_groups),
                                                       int(total_ctas), int(batch), int(k), _stream()), 0.0, _e)


# --------------------------------------------------------------- attention -------
def attention_fwd(q, k, v, out, batch, heads, nq, nk, ldq, ldk, ldv, ldo, lse=None, causal=False):
    a = AttentionArgs(_p(q), _p(k), _p(v), _p(out), _p(lse), batch, heads, nq, nk, 64, ldq, ldk, ldv, ldo, int(causal))
    _e = _ev()
    _run("attention_fwd", lib().ddpo_attention_fwd(C.byref(a), _stream()), 4.0 * batch * heads * nq * nk * 64, _e,
         f"B{batch} H{heads} nq{nq} nk{nk}")


# ------------------------------------------------------------------ wgrad -------
_WGRAD_WS = {}


def wgrad(*, dy, n, x0, dw, x1=None, c0=None, c1=0, ldy=0, ldx0=0, ldx1=0, conv=None, m=None, taps=1, stride=1, kernel=0):
    """dw[(tap, cin), n] += X^T dY (fp32, Flax layout).  conv=(batch, h_out, w_out) or linear with m rows."""
    from ._lib import WgradArgs
    a = WgradArgs()
    a.dy, a.ldy, a.n = _p(dy), int(ldy), int(n)
    a.x0, a.x1 = _p(x0), _p(x1)
    a.c0 = int(c0 if c0 is not None else x0.shape[-1])
    a.c1, a.ldx0, a.ldx1 = int(c1), int(ldx0), int(ldx1)
    if conv is not None:
        a.is_conv, a.batch, a.h, a.w, a.m = 1, int(conv[0]), int(conv[1]), int(conv[2]), 0
    else:
        a.is_conv, a.batch, a.h, a.w, a.m = 0, 0, 0, 0, int(m)
    a.conv_stride, a.taps = int(stride), int(taps)
    a.dw = _p(dw)
    a.kernel_override = int(kernel)
    need = int(lib().ddpo_wgrad_workspace_floats(C.byref(a)))
    dev = dw.device
    ws = _WGRAD_WS.get(dev)
    if need > 0 and (ws is None or ws.numel() < need):
        ws = torch.empty(max(need, 1 << 20), dtype=torch.float32, device=dev)
        _WGRAD_WS[dev] = ws
    a.workspace = _p(ws) if need > 0 else None
    a.workspace_floats = ws.numel() if (need > 0) else 0
    rows = a.batch * a.h * a.w if a.is_conv else a.m
    _e = _ev()
    _run("wgrad", lib().ddpo_wgrad(C.byref(a), _stream()), 2.0 * rows * a.n * a.taps * (a.c0 + a.c1), _e,
         f"M{rows} N{a.n} K{a.taps * (a.c0 + a.c1)} taps{a.taps} s{a.conv_stride}")


# ------------------------------------------------------- backward wrappers -------
def attention_bwd(q, k, v, out, dout, lse, delta, dq, dk, dv, batch, heads, nq, nk, ldq, ldk, ldv, ldo, lddo, lddq,
                  lddk, lddv):
    from ._lib import AttentionBwdArgs
    a = AttentionBwdArgs(_p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv), batch,
                         heads, nq, nk, 64, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, None)
    _e = _ev()
    _run("attention_bwd", lib().ddpo_attention_bwd(C.byref(a), _stream()), 14.0 * batch * heads * nq * nk * 64, _e,
         f"B{batch} H{heads} nq{nq} nk{nk}")


_CIW_WS = {}
_KERNELS_PER_CALL.update({"attention_bwd": 3, "colsum_cast": 2, "conv_out_bwd": 3, "conv_in_wgrad": 2,
                          "dense_small_bwd": 3, "grad_sumsq": 2, "wgrad": 2, "colsum_bf16": 2})
_COLSUM_WS = {}


def colsum_cast(dy, m, n, y_bf16=None, out=None, rows_per_group=None, accumulate=True, ld=0):
    """Optional bf16 copy of dy [m, n] and out[g, n] (+)= per-group column sums (bias / per-sample grads)."""
    rpg = int(rows_per_group or m)
    ws = None
    if out is not None:
        need = int(lib().ddpo_colsum_workspace_floats(m, n, rpg))
        ws = _COLSUM_WS.get(dy.device)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 1 << 20), dtype=torch.float32, device=dy.device)
            _COLSUM_WS[dy.device] = ws
    _e = _ev()
    _run("colsum_cast", lib().ddpo_colsum_cast(_p(dy), int(ld), _p(y_bf16), _p(out), rpg, int(accumulate), _p(ws), int(m),
                                               int(n), _stream()), float(m) * n * 6, _e)


def colsum_bf16(x, m, n, out, accumulate=True, ld=0):
    need = int(lib().ddpo_colsum_workspace_floats(m, n, m))
    ws = _COLSUM_WS.get(x.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), dtype=torch.float32, device=x.device)
        _COLSUM_WS[x.device] = ws
    _e = _ev()
    _run("colsum_bf16", lib().ddpo_colsum_bf16(_p(x), int(ld), _p(out), int(accumulate), _p(ws), int(m), int(n), _stream()),
         float(m) * n * 2, _e)


def geglu_bwd(pre, dff, dpre, m, n, bn=256):
    assert dff.dtype in (torch.float32, torch.bfloat16)
    _e = _ev()
    _run("geglu_bwd", lib().ddpo_geglu_bwd(_p(pre), _p(dff), _p(dpre), int(m), int(n), int(bn),
                                           int(dff.dtype == torch.bfloat16), _stream()), 0.0, _e)


def conv_out_bwd(x_nhwc, w, dy_nchw, dx_nhwc, dw, dbias, batch, h, wd, cin):
    need = int(lib().ddpo_conv_out_bwd_workspace_floats(cin))
    ws = _CIW_WS.get(("cow", dw.device))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=dw.device)
        _CIW_WS[("cow", dw.device)] = ws
    _e = _ev()
    _run("conv_out_bwd", lib().ddpo_conv_out_bwd(_p(x_nhwc), _p(w), _p(dy_nchw), _p(dx_nhwc), _p(dw), _p(dbias), _p(ws),
                                                 batch, h, wd, cin, _stream()), 0.0, _e)


def conv_in_wgrad(lat, dx_nhwc, dw, batch, cin, h, wd, cout):
    need = int(lib().ddpo_conv_in_wgrad_workspace_floats(cin, cout))
    ws = _CIW_WS.get(dw.device)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.float32, device=dw.device)
        _CIW_WS[dw.device] = ws
    _e = _ev()
    _run("conv_in_wgrad", lib().ddpo_conv_in_wgrad(_p(lat), _p(dx_nhwc), _p(dw), _p(ws), batch, cin, h, wd, cout,
                                                   _stream()), 0.0, _e)


def dense_small_bwd(x, w, bias, dy, dpre_ws, dw, db, dx, batch, k, n, silu_in=False, silu_out=False, dx_accumulate=False):
    _e = _ev()
    _run("dense_small_bwd", lib().ddpo_dense_small_bwd(_p(x), _p(w), _p(bias), _p(dy), _p(dpre_ws), _p(dw), _p(db), _p(dx),
                                                       int(dx_accumulate), batch, k, n, int(silu_in), int(silu_out),
                                                       _stream()), 0.0, _e)


def dilate2x_bf16(x, y, batch, h, w, c):
    _e = _ev()
    _run("dilate2x_bf16", lib().ddpo_dilate2x_bf16(_p(x), _p(y), batch, h, w, c, _stream()), 0.0, _e)


def copy2d(src, lds, dst, ldd, rows, cols, accumulate=False):
    _e = _ev()
    _run("copy2d", lib().ddpo_copy2d(_p(src), int(lds), _p(dst), int(ldd), int(rows), int(cols), int(accumulate),
                                     _stream()), 0.0, _e)


def gather_rows(src, index, dst):
    """dst[r] = src[index[r]] for fp32 [rows, row_floats] matrices; index int64 on the device."""
    _chk(src, torch.float32, "src")
    _chk(dst, torch.float32, "dst")
    _chk(index, torch.int64, "index")
    rows = index.numel()
    row_floats = dst.numel() // rows
    assert src.numel() % row_floats == 0 and dst.numel() == rows * row_floats
    _e = _ev()
    _run("gather_rows", lib().ddpo_gather_rows(_p(src), _p(index), _p(dst), rows, row_floats, _stream()),
         float(rows) * row_floats * 8, _e)


def optim_workspace(device):
    return torch.empty(int(lib().ddpo_optim_workspace_bytes()), dtype=torch.uint8, device=device)


def grad_sumsq(g, ws, out):
    _e = _ev()
    _run("grad_sumsq", lib().ddpo_grad_sumsq(_p(g), g.numel(), _p(ws), _p(out), _stream()), float(g.numel()) * 4, _e)


def clip_adamw(params, grad_acc, mu, nu, sumsq, grad_scale, max_norm, lr, b1, b2, eps, wd, step, norm_out=None):
    _e = _ev()
    _run("clip_adamw", lib().ddpo_clip_adamw(_p(params), _p(grad_acc), _p(mu), _p(nu), params.numel(), _p(sumsq),
                                             float(grad_scale), float(max_norm), float(lr), float(b1), float(b2),
                                             float(eps), float(wd), int(step), _p(norm_out), _stream()),
         float(params.numel()) * 24, _e)


# --------------------------------------------------------------------- RWR -------
def rwr_workspace(batch, device):
    return torch.zeros(int(lib().ddpo_rwr_workspace_floats(int(batch))), dtype=torch.float32, device=device)


def rwr_noisy_latents(moments_nhwc, key_sample_dev, key_noise_dev, timesteps, alphas_cumprod, noise_out, noisy_out,
                      latents_out=None, scaling=0.18215):
    """moments [B,h,w,2C] fp32 -> noise / noisy latents [B,C,h,w] (reference ddpo/training/diffusion.py:16-43)."""
    _chk(moments_nhwc, torch.float32, "moments")
    _chk(timesteps, torch.int32, "timesteps")
    _chk(noise_out, torch.float32, "noise_out")
    _chk(noisy_out, torch.float32, "noisy_out")
    b, h, w, c2 = moments_nhwc.shape
    assert timesteps.numel() == b and noise_out.numel() == b * (c2 // 2) * h * w == noisy_out.numel()
    _e = _ev()
    _run("rwr_noisy_latents", lib().ddpo_rwr_noisy_latents(_p(moments_nhwc), _p(key_sample_dev), _p(key_noise_dev),
                                                           _p(timesteps), _p(alphas_cumprod), float(scaling), b, c2 // 2,
                                                           h, w, _p(noise_out), _p(noisy_out), _p(latents_out),
                                                           _stream()), 0.0, _e)


def rwr_mse_loss(eps_u, eps_c, noise, guidance, loss_out, ws, weights=None, per_sample=None, d_eps_u=None, d_eps_c=None):
    """loss of ddpo/training/diffusion.py:77-90 and its gradient wrt both U-Net outputs; eps_* / noise are [B, n]."""
    for t, nm in ((eps_u, "eps_u"), (eps_c, "eps_c"), (noise, "noise")):
        _chk(t, torch.float32, nm)
    b = noise.shape[0]
    n = noise.numel() // b
    if weights is not None:
        _chk(weights, torch.float32, "weights")
        assert weights.numel() == b
    _e = _ev()
    _run("rwr_mse_loss", lib().ddpo_rwr_mse_loss(_p(eps_u), _p(eps_c), _p(noise), _p(weights), float(guidance), b, n,
                                                 _p(loss_out), _p(per_sample), _p(d_eps_u), _p(d_eps_c), _p(ws),
                                                 _stream()), 0.0, _e)


# --------------------------------------------------------------------- VAE -------
def vae_post_quant(latents, w, bias, out, scaling=0.18215):
    _chk(latents, torch.float32, "latents")
    _chk(out, torch.float32, "out")
    b, c, h, wd = latents.shape
    _e = _ev()
    _run("vae_post_quant", lib().ddpo_vae_post_quant(_p(latents), _p(w), _p(bias), float(scaling), b, c, h, wd, _p(out),
                                                     _stream()), 0.0, _e)


def softmax_rows(scores, probs_bf16, scale):
    """probs[r] = softmax(scale * scores[r]); scores fp32 [rows, n], probs bf16 [rows, n] (row-contiguous)."""
    _chk(scores, torch.float32, "scores")
    _chk(probs_bf16, torch.bfloat16, "probs")
    rows, n = scores.shape
    _e = _ev()
    _run("softmax_rows", lib().ddpo_softmax_rows(_p(scores), n, float(scale), _p(probs_bf16), n, rows, n, _stream()),
         float(rows) * n * 6, _e)


def image_to_uint8(img, out_u8):
    """``(img * 255).astype(uint8)`` (the reference's truncating cast, callbacks.py:181) on the device"""
    _chk(img, torch.float32, "images")
    _chk(out_u8, torch.uint8, "out")
    assert img.numel() == out_u8.numel() and img.numel() % 4 == 0
    _e = _ev()
    _run("image_to_uint8", lib().ddpo_image_to_uint8(_p(img), _p(out_u8), img.numel(), _stream()), 0.0, _e)


def vae_image_to_nchw(img_nhwc, out_nchw):
    _chk(img_nhwc, torch.float32, "images")
    _chk(out_nchw, torch.float32, "out")
    b, h, w, c = img_nhwc.shape
    assert c == 3 and tuple(out_nchw.shape) == (b, 3, h, w)
    _e = _ev()
    _run("vae_image_to_nchw", lib().ddpo_vae_image_to_nchw(_p(img_nhwc), _p(out_nchw), b, h, w, _stream()), 0.0, _e)


def vae_encoder_head(x_nhwc, w, bias, wq, bq, moments, batch, h, wd, cin):
    _chk(moments, torch.float32, "moments")
    _e = _ev()
    _run("vae_encoder_head", lib().ddpo_vae_encoder_head(_p(x_nhwc), _p(w), _p(bias), _p(wq), _p(bq), _p(moments), batch, h,
                                                         wd, cin, _stream()), 0.0, _e)


def vae_conv_out(x_nhwc, w, bias, batch, h, wd, cin, raw_nchw=None, img_nhwc=None):
    _e = _ev()
    _run("vae_conv_out", lib().ddpo_vae_conv_out(_p(x_nhwc), _p(w), _p(bias), _p(raw_nchw), _p(img_nhwc), batch, h, wd,
                                                 cin, _stream()), 0.0, _e)


# ------------------------------------------------------------ text encoder -------
def embed_tokens(ids, token_embedding, position_embedding, out, seq_len):
    _chk(ids, torch.int32, "ids")
    _chk(out, torch.float32, "out")
    rows, dim = out.shape
    _e = _ev()
    _run("embed_tokens", lib().ddpo_embed_tokens(_p(ids), _p(token_embedding), _p(position_embedding), _p(out), rows,
                                                 int(seq_len), dim, token_embedding.shape[0], _stream()), 0.0, _e)


def act_bf16(x, y_bf16, act):
    """act: "gelu" (erf form) or "quick_gelu"."""
    _chk(x, torch.float32, "x")
    _chk(y_bf16, torch.bfloat16, "y")
    _e = _ev()
    _run("act_bf16", lib().ddpo_act_bf16(_p(x), _p(y_bf16), x.numel(), {"gelu": 0, "quick_gelu": 1}[act], _stream()),
         float(x.numel()) * 6, _e)


def layernorm_f32(x, scale, bias, y, m, c, eps=1e-5):
    _e = _ev()
    _run("layernorm_f32", lib().ddpo_layernorm_f32(_p(x), _p(scale), _p(bias), _p(y), int(m), int(c), float(eps),
                                                   _stream()), float(m) * c * 8, _e)


# ------------------------------------------------------------ image tower --------
def patchify_bf16(img_nhwc, out_bf16, patch):
    _chk(img_nhwc, torch.float32, "img")
    _chk(out_bf16, torch.bfloat16, "out")
    b, s, _, _ = img_nhwc.shape
    _e = _ev()
    _run("patchify_bf16", lib().ddpo_patchify_bf16(_p(img_nhwc), _p(out_bf16), b, s, int(patch), out_bf16.shape[1],
                                                   _stream()), 0.0, _e)


def vit_tokens(patches, class_embedding, position_embedding, out, batch, n_patches, dim):
    _e = _ev()
    _run("vit_tokens", lib().ddpo_vit_tokens(_p(patches), _p(class_embedding), _p(position_embedding), _p(out), int(batch),
                                             int(n_patches), int(dim), _stream()), 0.0, _e)


def l2norm_rows(x, y):
    _chk(x, torch.float32, "x")
    m, c = x.shape
    _e = _ev()
    _run("l2norm_rows", lib().ddpo_l2norm_rows(_p(x), _p(y), m, c, _stream()), 0.0, _e)
