"""NumPy restatement of the reference optimizer chain
``optax.chain(clip_by_global_norm(max_grad_norm), adamw(..., mu_dtype=bfloat16))``
(reference ``pipeline/policy_gradient.py:130-150``; optax==0.1.5) and of
``AccumulatingTrainState.apply_gradients`` (``ddpo/training/policy_gradient.py:32-48``).
TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch

f32 = np.float32


def to_bf16_f32(x):
    """Round a float32 array to bfloat16 (round-to-nearest-even) and widen back."""
    return torch.from_numpy(np.ascontiguousarray(x, f32)).to(torch.bfloat16).to(torch.float32).numpy()


class AdamWState:
    def __init__(self, n):
        self.count = 0
        self.mu = np.zeros(n, f32)   # holds bf16-representable values
        self.nu = np.zeros(n, f32)


def global_norm(g):
    return np.sqrt(np.sum(np.asarray(g, np.float64) ** 2)).astype(f32)


def clip_adamw_update(params, grads, st, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4, max_norm=1.0):
    p = np.asarray(params, f32)
    g = np.asarray(grads, f32)
    gn = global_norm(g)
    if not (gn < f32(max_norm)):
        g = ((g / gn) * f32(max_norm)).astype(f32)
    mu = (f32(1 - b1) * g + f32(b1) * st.mu).astype(f32)
    nu = (f32(1 - b2) * (g * g) + f32(b2) * st.nu).astype(f32)
    st.count += 1
    mu_hat = mu / f32(1 - b1 ** st.count)
    nu_hat = nu / f32(1 - b2 ** st.count)
    upd = mu_hat / (np.sqrt(nu_hat) + f32(eps))
    upd = upd + f32(wd) * p
    upd = f32(-lr) * upd
    st.mu = to_bf16_f32(mu)
    st.nu = nu
    return (p + upd).astype(f32), gn


class AccumulatingTrainState:
    """grad_acc += g; on do_update: g~ = (grad_acc + g)/(n_acc+1) -> tx -> zero."""

    def __init__(self, params, **hp):
        self.params = np.asarray(params, f32).copy()
        self.grad_acc = np.zeros_like(self.params)
        self.n_acc = 0
        self.step = 0
        self.opt = AdamWState(self.params.size)
        self.hp = hp

    def apply_gradients(self, grads, do_update):
        g = np.asarray(grads, f32)
        if do_update:
            gm = ((self.grad_acc + g) / f32(self.n_acc + 1)).astype(f32)
            self.params, gn = clip_adamw_update(self.params, gm, self.opt, **self.hp)
            self.grad_acc[:] = 0
            self.n_acc = 0
            self.step += 1
            return gn
        self.grad_acc = (self.grad_acc + g).astype(f32)
        self.n_acc += 1
        return None
