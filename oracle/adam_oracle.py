"""oracle/adam_oracle.py — TEST INFRASTRUCTURE (parity checker; never imported by the product).

numpy fp32 restatement of the optimiser step the reference takes for every parameter group of every GaussianModel:
`torch.optim.Adam(l, lr=0.0, eps=1e-15)` (/root/reference/lib/scene/gaussian_model.py:201), i.e. PyTorch's single-tensor Adam
(`torch/optim/adam.py::_single_tensor_adam`, amsgrad=False, weight_decay=0, maximize=False; the reference pins torch 2.3.1,
docs/INSTALL.md:32 — the algorithm is unchanged in the 2.11 of this image, against which tests/test_optim.py pins this file):

    exp_avg.lerp_(grad, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step;  bias_correction2 = 1 - beta2 ** step          (python floats)
    step_size = lr / bias_correction1;     bias_correction2_sqrt = bias_correction2 ** 0.5
    denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-step_size)
"""
import numpy as np

f32 = np.float32


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """One step on fp32 arrays; returns new (p, m, v). `step` is the 1-based count of this update."""
    p, g, m, v = (np.asarray(a, f32) for a in (p, g, m, v))
    w1, b2, w2 = f32(1.0 - beta1), f32(beta2), f32(1.0 - beta2)
    m = m + (g - m) * w1                                  # lerp, |weight| < 0.5 branch
    v = v * b2 + (w2 * g) * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    neg_step = f32(-(lr / bc1))
    den = np.sqrt(v) / f32(bc2 ** 0.5) + f32(eps)
    p = p + (neg_step * m) / den
    return p.astype(f32), m.astype(f32), v.astype(f32)
