#!/usr/bin/env python
"""One small hot-path step on cuda:0; prints how many sublayer backward passes received the bf16 operand copy written by
the LayerNorm backward (ops.grad_operand) and how many fell back to the fused cast."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtos_b200 import hotpath, ops, synthetic          # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    for p in (0.0, 0.2):
        cfg = hotpath.HotPathConfig(embed_dim=128, ff_embed_dim=256, num_heads=8, graph_layers=2, snt_layers=1,
                                    inference_layers=1, rnn_hidden_size=64, dropout=p, vocab_size=500)
        model = hotpath.HotPath(cfg).to(dev)
        model.train(p > 0)
        g = synthetic.make_batch(8, 16, 128, T_max=12, T_min=6, V=500, seed=3)
        batch = {k: v.to(dev) for k, v in hotpath.batch_tensors(g).items()}
        for k in ops.stats:
            ops.stats[k] = 0
        model(batch).backward()
        torch.cuda.synchronize()
        print(f"dropout {p}: {ops.stats}")


if __name__ == "__main__":
    main()
