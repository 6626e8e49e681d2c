"""NumPy restatement of JAX's threefry2x32 PRNG (jax==0.4.8, ``jax/_src/prng.py``
and ``jax/_src/random.py``; pinned at reference ``environment_tpu.yml:17-24``).

Call sites on the hot path (reference file:line):
  * ``pipeline/policy_gradient.py:51,201,244-245``  PRNGKey / split lineage
  * ``ddpo/diffusers_patch/pipeline_flax_stable_diffusion.py:196-197,232,252``
  * ``ddpo/diffusers_patch/scheduling_ddim_flax.py:347``  ``jax.random.normal``

Everything is uint32 / float32, evaluated without FMA contraction.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32_pair(k0, k1, x0, x1):
    """20-round Threefry-2x32 on arrays of counter words (x0, x1)."""
    with np.errstate(over="ignore"):
        k0 = np.uint32(k0)
        k1 = np.uint32(k1)
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = (np.asarray(x0, np.uint32) + ks[0]).astype(np.uint32)
        x1 = (np.asarray(x1, np.uint32) + ks[1]).astype(np.uint32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(i + 2) % 3] + np.uint32(i + 1)).astype(np.uint32)
    return x0, x1


def threefry_2x32(key, count):
    """``jax._src.prng.threefry_2x32``: counters are split in two halves
    (first half -> word 0, second half -> word 1); odd sizes are zero padded."""
    count = np.asarray(count, np.uint32)
    flat = count.ravel()
    odd = flat.size % 2
    if odd:
        flat = np.concatenate([flat, np.zeros(1, np.uint32)])
    h = flat.size // 2
    y0, y1 = threefry2x32_pair(key[0], key[1], flat[:h], flat[h:])
    out = np.concatenate([y0, y1])
    if odd:
        out = out[:-1]
    return out.reshape(count.shape)


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], np.uint32)


def split(key, num=2):
    counts = np.arange(num * 2, dtype=np.uint32)
    return threefry_2x32(key, counts).reshape(num, 2)


def random_bits(key, shape):
    size = int(np.prod(shape)) if len(shape) else 1
    return threefry_2x32(key, np.arange(size, dtype=np.uint32)).reshape(shape)


def uniform(key, shape, minval=0.0, maxval=1.0):
    bits = random_bits(key, shape)
    fb = (bits >> np.uint32(9)) | np.uint32(0x3F800000)
    floats = fb.view(np.float32) - np.float32(1.0)
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, floats * (maxval - minval) + minval).astype(np.float32)


def erfinv_f32(x):
    """XLA's float32 ErfInv (Giles' single-precision polynomial)."""
    x = np.asarray(x, np.float32)
    one = np.float32(1.0)
    w = -np.log1p(-(x * x)).astype(np.float32)
    lt = w < np.float32(5.0)
    ws = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
    c_lt = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
            0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
    c_ge = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
            0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]
    p = np.where(lt, np.float32(c_lt[0]), np.float32(c_ge[0])).astype(np.float32)
    for a, b in zip(c_lt[1:], c_ge[1:]):
        p = (np.where(lt, np.float32(a), np.float32(b)) + p * ws).astype(np.float32)
    r = (p * x).astype(np.float32)
    return np.where(np.abs(x) == one, np.float32(np.inf) * x, r).astype(np.float32)


def normal(key, shape):
    """``jax.random.normal`` (float32): sqrt(2) * erfinv(uniform(nextafter(-1,0), 1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = uniform(key, shape, lo, 1.0)
    return (np.float32(np.sqrt(2)) * erfinv_f32(u)).astype(np.float32)


def randint(key, shape, minval, maxval):
    """``jax.random.randint`` for int32 (jax==0.4.8 ``jax/_src/random.py:_randint``): two independent 32-bit draws
    ``higher, lower`` from ``split(key)``; ``offset = ((higher % span) * (2**32 % span) + lower % span) % span`` in
    uint32 arithmetic, ``2**32 % span`` formed as ``((2**16 % span) ** 2) % span``.  Call site: reference
    ``ddpo/training/diffusion.py:29-34`` (timesteps).  Unpinned against a live JAX run (none available offline)."""
    k1, k2 = split(key)
    higher = random_bits(k1, shape)
    lower = random_bits(k2, shape)
    span = np.uint32(max(int(maxval) - int(minval), 1))
    with np.errstate(over="ignore"):
        mult = np.uint32(2 ** 16) % span
        mult = np.uint32((mult * mult) % span)
        off = ((higher % span) * mult + (lower % span)).astype(np.uint32) % span
    return (np.int32(minval) + off.astype(np.int32)).astype(np.int32)
