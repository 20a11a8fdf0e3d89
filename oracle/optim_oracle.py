"""CPU oracle for the optimizer step of the reference's training loop.

TEST INFRASTRUCTURE ONLY (see oracle/gtos_oracle.py for the rule): nothing under ``gtos_b200/`` imports this.

Restates, per parameter and in fp32 torch CPU ops,
  clip_grad_norm_(params, max_norm)            generator/train.py:152 (torch.nn.utils: total 2-norm over all gradients,
                                               coef = max_norm / (total + 1e-6), applied only when < 1)
  AdamWeightDecayOptimizer.step()              generator/adam.py:28-87 (no bias correction; decoupled weight decay
                                               folded into the update: p -= lr * (m / (sqrt(v) + eps) + wd * p))
  the two parameter groups                     generator/train.py:123-132 (wd on non-bias, non-LayerNorm parameters)
  update_lr                                    generator/train.py:81-83

Parity pin: tests/golden/make_golden_optim.py runs the reference's own AdamWeightDecayOptimizer + torch's
clip_grad_norm_ for three steps on seeded parameters / gradients; tests/test_oracle_golden.py holds this file to
those vectors at 1e-6.
"""
import torch


def no_decay(name):
    """train.py:126"""
    return name.endswith("bias") or "layer_norm" in name


def update_lr(embed_size, steps, warmup_steps):
    """train.py:81-83"""
    return embed_size ** -0.5 * min(steps ** -0.5, steps * (warmup_steps ** -1.5))


def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (norm_type 2): returns (total_norm, coefficient actually applied)"""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = max_norm / (total + 1e-6)
    return total, torch.clamp(coef, max=1.0)


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, weight_decays, betas=(0.9, 0.999), eps=1e-6, max_norm=1.0):
    """One training-loop update, in place on the lists `params`, `exp_avg`, `exp_avg_sq` (dict name -> tensor).
    grads: dict name -> tensor (not modified); weight_decays: dict name -> float."""
    b1, b2 = betas
    names = list(params)
    coef = 1.0
    if max_norm is not None:
        _, coef = clip_coef([grads[n] for n in names], max_norm)
    for n in names:
        g = grads[n] * coef
        exp_avg[n].mul_(b1).add_(g, alpha=1 - b1)                       # adam.py:69
        exp_avg_sq[n].mul_(b2).addcmul_(g, g, value=1 - b2)             # adam.py:70
        denom = exp_avg_sq[n].sqrt().add_(eps)                          # adam.py:77
        update = (exp_avg[n] / denom).add_(params[n], alpha=weight_decays[n])   # adam.py:86
        params[n].add_(update, alpha=-lr)                               # adam.py:87
