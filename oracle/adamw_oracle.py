"""CPU restatement of the optimizer on the reference's training path — TEST INFRASTRUCTURE ONLY (imported by tests/ only).

The reference steps `pytorch_transformers.AdamW` (requirements.txt pins pytorch-transformers==1.0.0; the package is a
third-party dependency that is NOT vendored under /root/reference), built at train_tasks.py:401-426 with one param group per
tensor (lr 1e-4 for `vil_*` heads, base_lr otherwise; weight_decay 0.0 for bias / LayerNorm, 0.01 otherwise) and
correct_bias=False, and stepped at train_tasks.py:550. Its published algorithm (pytorch_transformers/optimization.py, class
AdamW.step, v1.0.0) is restated here:

    state.step += 1
    exp_avg    = beta1 * exp_avg    + (1 - beta1) * grad
    exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * grad * grad
    denom      = sqrt(exp_avg_sq) + eps
    step_size  = lr                                   (correct_bias False)
               = lr * sqrt(1 - beta2^t) / (1 - beta1^t)   (correct_bias True)
    p         -= step_size * exp_avg / denom
    p         -= lr * weight_decay * p                (only if weight_decay > 0; AFTER the Adam update, on the updated p)

Parity pinning: the package itself is absent, so this restatement is anchored on the reference's call sites above and on
torch.optim.AdamW as an independent cross-check where the two algorithms coincide (correct_bias=True and weight_decay == 0:
identical up to where eps enters; see tests/test_optim.py).
"""
import math

import torch


def adamw_step(p, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0, correct_bias=True):
    """One in-place AdamW update of a single tensor; `step` is the 1-based step count after the increment."""
    exp_avg.mul_(beta1).add_(grad, alpha=1.0 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
    denom = exp_avg_sq.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(exp_avg, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def reference_param_groups(named_parameters, base_lr, vision_scratch=False):
    """The grouping of train_tasks.py:401-421 (freeze == -1): lr 1e-4 for names containing 'vil_', base_lr otherwise; no weight
    decay for names containing 'bias', 'LayerNorm.bias' or 'LayerNorm.weight', 0.01 otherwise. One group per tensor."""
    no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
    groups = []
    for key, value in named_parameters:
        if not value.requires_grad:
            continue
        lr = 1e-4 if "vil_" in key else base_lr
        wd = 0.0 if any(nd in key for nd in no_decay) else 0.01
        groups.append({"params": [value], "lr": lr, "weight_decay": wd})
    return groups
