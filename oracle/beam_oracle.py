"""CPU oracle for the beam search bookkeeping of the reference, on token IDS.

TEST INFRASTRUCTURE ONLY: nothing under ``gtos_b200/`` imports this.

Restates generator/search.py in plain Python lists:
  Hypothesis                 search.py:11-30   (seq starts with <STR>; completed when seq[-1] == END)
  Beam.merge_score           search.py:42-55   (UNK -> -inf, else prev score + token score)
  Beam.update                search.py:57-92   (all (hyp, token) candidates, stable descending sort, keep
                                                beam_size - #completed, END goes to `completed` only when
                                                len(seq) - 2 >= min_time_step, otherwise the hypothesis is dropped)
  Beam.completed / get_k_best  search.py:93-102
  search_by_batch            search.py:114-168 (beams advance in lock step; finished beams submit nothing)
The reference works on token strings and hidden-state dicts; here a hypothesis is (seq of ids, score) and the model is a
callback, which is all the bookkeeping depends on.

Parity pin: tests/golden/make_golden_beam.py drives the reference's own Beam / search_by_batch with a scripted
decode_step; tests/test_beam_cpu.py checks this file (and gtos_b200.decode.BeamState) against those runs.
"""


class Hyp:
    def __init__(self, seq, score, parent=None):
        self.seq, self.score, self.parent = seq, score, parent


class BeamOracle:
    def __init__(self, beam_size, min_time_step, max_time_step, start_id, end_id, unk_id):
        self.K, self.Tmin, self.Tmax = beam_size, min_time_step, max_time_step
        self.end_id, self.unk_id = end_id, unk_id
        self.hyps = [Hyp([start_id], 0.0)]
        self.done = []
        self.steps = 0

    def completed(self):
        return not (len(self.done) < self.K and self.steps < self.Tmax)

    def update(self, last_steps):
        """last_steps: per live hypothesis, list of (token id, score) best first (search.py:58)"""
        cands = []
        for i, steps in enumerate(last_steps):
            for tok, sc in steps:
                s = float("-inf") if tok == self.unk_id else self.hyps[i].score + sc
                cands.append((i, tok, s))
        cands.sort(key=lambda c: c[2], reverse=True)
        cands = cands[:self.K - len(self.done)]
        new = [Hyp(self.hyps[i].seq + [tok], s, parent=i) for i, tok, s in cands]
        self.hyps = []
        for h in new:
            if h.seq[-1] == self.end_id:
                if len(h.seq) - 2 >= self.Tmin:
                    self.done.append(h)
            else:
                self.hyps.append(h)
        self.steps += 1

    def k_best(self, k, alpha):
        if not self.done:
            self.done = self.hyps
        self.done.sort(key=lambda h: h.score / ((1 + len(h.seq)) ** alpha), reverse=True)
        return self.done[:k]


def search_by_batch(beams, step_fn, topk):
    """step_fn(list of (beam index, Hyp), t) -> list (same order) of [(token id, score)] * topk, best first.
    Mirrors search.py:114-168: every live hypothesis of every unfinished beam is submitted in beam order."""
    t = 0
    while True:
        subs = [(b, h) for b, beam in enumerate(beams) if not beam.completed() for h in beam.hyps]
        if not subs:
            break
        results = step_fn(subs, t)
        pos = 0
        for beam in beams:
            if not beam.completed():
                n = len(beam.hyps)
                beam.update(results[pos:pos + n])
                pos += n
        t += 1
    return beams
