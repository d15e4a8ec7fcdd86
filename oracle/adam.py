""" TEST INFRASTRUCTURE ONLY — numpy (fp32) restatement of the optimizer step of the fit loop.

The reference builds `torch.optim.Adam(params, lr=lr)` on every `fit` (pydens/model_torch.py:419-422) and calls
`optimizer.step()` once per iteration (:461).  torch is a third-party dependency of the reference (present here:
2.11); its published algorithm (Kingma & Ba 2015, as implemented by `torch/optim/adam.py`, no AMSGrad, L2 weight decay
added to the gradient) is, per element, with t = 1, 2, ...:

    g  <- g + weight_decay * p
    m  <- m + (1 - beta1) * (g - m)                       (lerp form)
    v  <- beta2 * v + (1 - beta2) * g * g
    p  <- p - lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)

This is the arithmetic the step kernel's tail applies (pydens_b200/csrc/pinn_step_kernel.cuh: adam_hyper / adam_apply)
and the persistent kernels apply in their loop.  Pinned against torch.optim.Adam itself (tests/test_oracle.py); the GPU
tests compare the kernels with torch's Adam on the device.
"""
import numpy as np


def adam_step(p, g, m, v, t, lr=0.005, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, mask=None):
    """ One Adam step on flat fp32 arrays, in place; `t` is the 1-based step number, `mask` (optional) marks the
    trainable entries (frozen ones are left untouched, like parameters outside the optimizer). """
    f = np.float32
    p, g, m, v = (np.asarray(a, dtype=np.float32) for a in (p, g, m, v))
    sel = slice(None) if mask is None else np.asarray(mask) != 0
    gg = g[sel]
    if weight_decay != 0.0:
        gg = gg + f(weight_decay) * p[sel]
    m1 = m[sel] + f(1.0 - beta1) * (gg - m[sel])
    v1 = f(beta2) * v[sel] + f(1.0 - beta2) * gg * gg
    step_size = f(lr) / f(1.0 - np.power(f(beta1), f(t)))
    bc2_sqrt = np.sqrt(f(1.0 - np.power(f(beta2), f(t))), dtype=np.float32)
    p[sel] = p[sel] - step_size * m1 / (np.sqrt(v1) / bc2_sqrt + f(eps))
    m[sel] = m1
    v[sel] = v1
    return p, m, v
