"""Golden runs of the reference's OWN beam search bookkeeping (generator/search.py: Hypothesis, Beam, search_by_batch)
driven by a scripted decode_step (tests/beam_script.py), so the merge / completion / ranking rules are pinned without a
model.  Build container only (needs /root/reference):
    python tests/golden/make_golden_beam.py        -> tests/golden/golden_beam_v1.json

Shim (SURVEY.md §8c style, no effect on the logic): Tensor.cuda is made a no-op because search.py moves its index
tensors with `.cuda(device)` (search.py:72,136) and this container has no GPU.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/generator")
import beam_script as BS                          # noqa: E402
import search as ref_search                       # noqa: E402
from data import END, UNK, STR                    # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self


def tok2str(i):
    return {BS.END: END, BS.UNK: UNK, BS.START: STR}.get(i, f"w{i}")


def str2tok(s):
    return {END: BS.END, UNK: BS.UNK, STR: BS.START}.get(s) if s in (END, UNK, STR) else int(s[1:])


class ScriptedModel:
    """the two methods search_by_batch needs (search.py:4-8)"""

    def __init__(self, case, W):
        self.case, self.W = case, W

    def prepare_incremental_input(self, step_seq):
        return [s[0] for s in step_seq]

    def decode_step(self, inp, state_dict, mem_dict, offset, topk):
        last = [str2tok(s) for s in inp]
        n = len(last)
        col = torch.tensor(last, dtype=torch.float32).view(1, n, 1)
        prefix = torch.cat([state_dict["prefix"], col], 0) if "prefix" in state_dict else col      # [t+1, Hyp, 1]
        results = []
        for h in range(n):
            seq = [int(v) for v in prefix[:, h, 0].tolist()]
            row = torch.from_numpy(BS.table(self.case, mem_dict["src"][h], seq, self.W))
            sc, ix = torch.topk(row, topk)
            results.append([(tok2str(int(i)), float(s)) for s, i in zip(sc, ix)])
        return {"prefix": prefix}, results


def main():
    out = {}
    for name, c in BS.CASES:
        beams = [ref_search.Beam(c["K"], c["Tmin"], c["Tmax"], [ref_search.Hypothesis({}, [STR], 0.)], "cpu")
                 for _ in range(c["B"])]
        ref_search.search_by_batch(ScriptedModel(name, c["W"]), beams, {"src": list(range(c["B"]))})
        res = []
        for beam in beams:
            best = beam.get_k_best(c["K"], c["alpha"])
            res.append(dict(steps=beam.steps, hyps=[dict(seq=[str2tok(s) for s in h.seq], score=h.score) for h in best]))
        out[name] = res
    with open(os.path.join(HERE, "golden_beam_v1.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: [len(b["hyps"]) for b in v] for k, v in out.items()})


if __name__ == "__main__":
    main()
