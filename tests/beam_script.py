"""Scripted log-prob tables for the beam-search bookkeeping tests: a deterministic function of (source graph, prefix),
shared by tests/golden/make_golden_beam.py (which feeds it to the reference's own Beam / search_by_batch) and by the tests
that drive the oracle restatement and gtos_b200.decode.BeamState.  numpy's legacy RandomState is stable across versions."""
import zlib

import numpy as np

START, END, UNK = 2, 3, 1          # ids used by the scripted vocabulary (0 = pad)


def table(case, b, seq, W):
    """log-prob row for the hypothesis of graph `b` whose token ids so far are `seq` (incl. START)."""
    rs = np.random.RandomState(zlib.crc32(repr((case, b, tuple(seq))).encode()))
    logits = rs.randn(W).astype(np.float32) * 2.0
    logits[END] += 1.5 * (len(seq) - 2)              # END becomes likely as the prefix grows
    logits[UNK] += 1.0                               # UNK is often in the top-k (scored -inf by merge_score)
    m = logits.max()
    return (logits - (m + np.log(np.exp(logits - m).sum()))).astype(np.float32)


CASES = [
    # name: (B graphs, beam K, vocabulary width W, min_time_step, max_time_step, alpha)
    ("small", dict(B=5, K=3, W=11, Tmin=1, Tmax=7, alpha=0.6)),
    ("wide", dict(B=4, K=8, W=40, Tmin=2, Tmax=10, alpha=0.6)),
    ("maxlen", dict(B=3, K=4, W=30, Tmin=50, Tmax=6, alpha=1.0)),     # nothing can complete: falls back to live hyps
    ("beam1", dict(B=4, K=1, W=9, Tmin=1, Tmax=8, alpha=0.0)),
]
