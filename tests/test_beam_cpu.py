"""Beam-search bookkeeping, CPU only.

  * oracle/beam_oracle.py (restatement of generator/search.py on token ids) against golden runs of the reference's own
    Beam / search_by_batch (tests/golden/make_golden_beam.py);
  * gtos_b200.decode.BeamState (the fixed-shape, sync-free device implementation; pure index arithmetic, so it runs on CPU
    tensors too) against the same golden runs: identical sequences, scores to 1e-5.
"""
import json
import os

import pytest
import torch

import beam_script as BS
from oracle import beam_oracle as BO

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_beam_v1.json")))


def _topk(row, k):
    sc, ix = torch.topk(torch.from_numpy(row), k)
    return [(int(i), float(s)) for s, i in zip(sc, ix)]


@pytest.mark.parametrize("name,c", BS.CASES)
def test_beam_oracle_matches_reference_search(name, c):
    beams = [BO.BeamOracle(c["K"], c["Tmin"], c["Tmax"], BS.START, BS.END, BS.UNK) for _ in range(c["B"])]

    def step_fn(subs, t):
        return [_topk(BS.table(name, b, h.seq, c["W"]), c["K"]) for b, h in subs]

    BO.search_by_batch(beams, step_fn, c["K"])
    for beam, ref in zip(beams, GOLD[name]):
        best = beam.k_best(c["K"], c["alpha"])
        assert beam.steps == ref["steps"]
        assert [h.seq for h in best] == [r["seq"] for r in ref["hyps"]]
        for h, r in zip(best, ref["hyps"]):
            assert h.score == pytest.approx(r["score"], abs=1e-5) or (h.score == r["score"] == float("-inf"))


@pytest.mark.parametrize("name,c", BS.CASES)
def test_beam_state_matches_reference_search(name, c):
    from gtos_b200.decode import BeamState
    B, K, W = c["B"], c["K"], c["W"]
    st = BeamState(B, K, c["Tmax"], c["Tmin"], BS.END, BS.UNK, torch.device("cpu"))
    for t in range(c["Tmax"]):
        if not bool(st.active().any()):
            break
        table = torch.zeros(B * K, W)
        for b in range(B):
            for s in range(K):
                if bool(st.live[b, s]) and bool(st.active()[b]):
                    seq = [BS.START] + (st._trace(st.tok, st.par, t - 1, b, s) if t > 0 else [])
                    table[b * K + s] = torch.from_numpy(BS.table(name, b, seq, W))
        st.update(t, table)
    got = st.k_best(K, c["alpha"])
    for b, ref in enumerate(GOLD[name]):
        assert int(st.steps[b]) == ref["steps"]
        assert [[BS.START] + seq for seq, _ in got[b]] == [r["seq"] for r in ref["hyps"]], (name, b)
        for (_, score), r in zip(got[b], ref["hyps"]):
            assert score == pytest.approx(r["score"], abs=1e-5) or (score == r["score"] == float("-inf"))


def test_dead_rows_never_outrank_real_unk_candidates():
    """a real candidate scored -inf (UNK, search.py:51-52) must still be taken before anything from a dead slot"""
    from gtos_b200.decode import BeamState
    st = BeamState(1, 2, 4, 1, BS.END, BS.UNK, torch.device("cpu"))
    table = torch.full((2, 6), -10.0)
    table[0, BS.UNK] = -0.1
    table[0, 4] = -0.5
    table[1] = 0.0                                   # dead slot with attractive scores
    st.update(0, table)
    assert st.live.tolist() == [[True, True]]
    assert st.tok[0, 0].tolist() == [4, BS.UNK] and st.score[0, 1] == float("-inf")
