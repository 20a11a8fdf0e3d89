"""Golden vectors for the optimizer step from the reference's OWN AdamWeightDecayOptimizer
(generator/adam.py) + torch.nn.utils.clip_grad_norm_ + update_lr (generator/train.py:81-83,123-132,151-153).
Build container only (needs /root/reference):
    python tests/golden/make_golden_optim.py       -> tests/golden/golden_optim_v1.pt
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/generator")
from adam import AdamWeightDecayOptimizer        # noqa: E402

warnings.filterwarnings("ignore")                # adam.py uses the deprecated add_(Number, Tensor) overloads
SEED = 19940117
SHAPES = {"enc.layers.0.fc1.weight": (37, 16), "enc.layers.0.fc1.bias": (37,), "enc.layers.0.attn_layer_norm.weight": (16,),
          "enc.layers.0.attn_layer_norm.bias": (16,), "enc.in_proj_weight": (48, 16), "enc.in_proj_bias": (48,),
          "embed.weight": (101, 7), "rnn.weight_hh_l0": (24, 8), "rnn.bias_hh_l0": (24,)}


def update_lr(optimizer, embed_size, steps, warmup_steps):
    # train.py:81-83 verbatim semantics (train.py itself cannot be imported without the data files' argparse flow)
    for g in optimizer.param_groups:
        g["lr"] = embed_size ** -0.5 * min(steps ** -0.5, steps * (warmup_steps ** -1.5))


def main():
    gen = torch.Generator().manual_seed(SEED)
    params = {n: torch.nn.Parameter(torch.randn(s, generator=gen) * 0.5) for n, s in SHAPES.items()}
    init = {n: p.detach().clone() for n, p in params.items()}
    decay = [p for n, p in params.items() if not (n.endswith("bias") or "layer_norm" in n)]      # train.py:123-132
    rest = [p for n, p in params.items() if n.endswith("bias") or "layer_norm" in n]
    opt = AdamWeightDecayOptimizer([{"params": decay, "weight_decay": 1e-4}, {"params": rest, "weight_decay": 0.}],
                                   lr=1e-3, betas=(0.9, 0.999), eps=1e-6)
    steps = []
    for k, scale in enumerate([3.0, 0.02, 1.0, 0.3], start=1):          # clipped, not clipped, clipped, borderline
        grads = {n: torch.randn(s, generator=gen) * scale / 10 for n, s in SHAPES.items()}
        for n, p in params.items():
            p.grad = grads[n].clone()
        total = torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)                        # train.py:152
        update_lr(opt, 512, k, 3)
        opt.step()
        steps.append(dict(grads=grads, lr=opt.param_groups[0]["lr"], total_norm=float(total),
                          params={n: p.detach().clone() for n, p in params.items()},
                          exp_avg={n: opt.state[p]["exp_avg"].clone() for n, p in params.items()},
                          exp_avg_sq={n: opt.state[p]["exp_avg_sq"].clone() for n, p in params.items()}))
    torch.save(dict(init=init, steps=steps, embed_size=512, warmup=3), os.path.join(HERE, "golden_optim_v1.pt"))
    print("total norms", [s["total_norm"] for s in steps], "lrs", [s["lr"] for s in steps])


if __name__ == "__main__":
    main()
