"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's LAMB step.

Follows /root/reference/peract/helpers/optim/lamb.py:60-122 operation for operation (no bias correction,
weight decay added to the Adam step, trust ratio = clamp(|w|, 0, 10) / |step|, 1 when either norm is zero).
Pinned by tests/test_train_ops.py against the reference class itself when /root/reference is mounted.
Nothing in voxactb_b200/ imports this module."""
import torch


def lamb_step(p, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay):
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)                       # lamb.py:95
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)          # :97
    weight_norm = p.pow(2).sum().sqrt().clamp(0, 10)                      # :105
    adam_step = exp_avg / exp_avg_sq.sqrt().add(eps)                      # :107
    if weight_decay != 0:
        adam_step.add_(p, alpha=weight_decay)                             # :109
    adam_norm = adam_step.pow(2).sum().sqrt()                             # :111
    trust = 1 if (weight_norm == 0 or adam_norm == 0) else weight_norm / adam_norm   # :112-115
    p.add_(adam_step, alpha=-lr * trust)                                  # :122 (fp32 tensor arithmetic when trust is a tensor)
    return p
