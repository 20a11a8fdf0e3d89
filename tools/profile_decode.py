#!/usr/bin/env python
"""Run a few eager decode steps of the cfg5-sized engine between cudaProfilerStart/Stop (for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`).  Usage: python tools/profile_decode.py [t0] [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gtos_b200 import hotpath                                   # noqa: E402
from gtos_b200.decode import BeamSearchDevice, DecodeEngine     # noqa: E402


def main():
    t0 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    torch.manual_seed(19940117)
    cfg = hotpath.HotPathConfig()
    model = hotpath.HotPath(cfg).to(dev).eval()
    B, K, S, D, V, steps = 256, 8, 40, 512, 10000, 32
    gen = torch.Generator().manual_seed(19940117)
    graph = torch.randn(S, B, D, generator=gen).to(dev)
    lens = torch.randint(S // 2, S + 1, (B,), generator=gen)
    gmask = (torch.arange(S).unsqueeze(1) >= lens.unsqueeze(0)).to(dev)
    probe = torch.tanh(torch.randn(1, B, D, generator=gen)).to(dev)
    copy_seq = torch.randint(2, V + 16, (S, B), generator=gen).to(dev)
    Wt = V + 16
    emb = torch.randn(Wt, D, generator=gen).to(dev)
    pos = torch.randn(steps, D, generator=gen).to(dev)
    eng = DecodeEngine(model.snt_encoder, model.decoder, max_hyp=B * K, max_steps=steps)
    eng.set_memory(graph, gmask, probe, copy_seq, table_width=Wt)
    bs = BeamSearchDevice(eng, K, steps, 1, 3, 1, 2, lambda tok, t: torch.nn.functional.layer_norm(emb[tok] + pos[t], (D,)))
    bs._alloc(B)
    bs.reset()
    for t in range(t0):
        bs._step(t)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for t in range(t0, t0 + n):
        bs._step(t)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
