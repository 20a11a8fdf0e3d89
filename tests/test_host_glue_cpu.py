"""CPU: host-side construction helpers of the drop-in encoder module (reference: generator/encoder.py:9-64)."""
import torch

from gtos_b200.encoder import AMREmbedding, RelationEncoder


class Vocab:
    def __init__(self, toks):
        self.toks, self.size, self.padding_idx, self.unk_idx = toks, len(toks), 0, 1

    def idx2token(self, i):
        return self.toks[i]


def test_pretrained_embedding_rows_statistics_and_dump(tmp_path):
    toks = ["<PAD>", "<UNK>", "want-01", "boy", "go-02", "unseen", "girl"]
    vec = {"want": [1.0, 2.0, 3.0, 4.0], "boy": [0.5, -0.5, 0.25, -0.25], "go-02": [9.0, 8.0, 7.0, 6.0], "girl": [0.0, 1.0, 0.0, 1.0],
           "other": [5.0, 5.0, 5.0, 5.0]}
    f = tmp_path / "vectors.txt"
    lines = [f"{k} " + " ".join(str(x) for x in v) for k, v in vec.items()] + ["bad 1.0 2.0"]
    f.write_text("\n".join(lines) + "\n", encoding="utf8")
    dump = tmp_path / "dump.txt"
    torch.manual_seed(0)
    emb = AMREmbedding(Vocab(toks), 4, pretrained_file=str(f), amr=True, dump_file=str(dump))
    W = emb.weight.detach()
    assert emb.weight.requires_grad and W.shape == (7, 4)
    assert torch.equal(W[0], torch.zeros(4))                                     # padding row
    assert torch.equal(W[2], torch.tensor(vec["want"]))                          # sense suffix ignored with amr=True
    assert torch.equal(W[3], torch.tensor(vec["boy"])) and torch.equal(W[6], torch.tensor(vec["girl"]))
    assert torch.isfinite(W[4]).all() and not torch.equal(W[4], torch.tensor(vec["go-02"]))   # file line `go-02` is not kept
    assert torch.isfinite(W[5]).all() and not torch.equal(W[5], torch.zeros(4))  # unseen token: random draw
    kept = dump.read_text(encoding="utf8").splitlines()
    assert [l.split(" ")[0] for l in kept] == ["want", "boy", "girl"]            # `go-02` normalises to `go`: not in the file
    # amr=False: the literal token has to match
    emb2 = AMREmbedding(Vocab(toks), 4, pretrained_file=str(f), amr=False)
    assert torch.equal(emb2.weight[4].detach(), torch.tensor(vec["go-02"]))
    assert not torch.equal(emb2.weight[2].detach(), torch.tensor(vec["want"]))
    # no file: N(0, 0.02) table with a zero padding row
    emb3 = AMREmbedding(Vocab(toks), 4)
    assert torch.equal(emb3.weight[0].detach(), torch.zeros(4)) and emb3.weight.detach().abs().max() < 0.2


def test_relation_encoder_constructor_variants():
    class V:
        size, padding_idx, unk_idx = 17, 0, 1

    bi = RelationEncoder(V(), 12, 32, 16, 2, 0.1)
    uni = RelationEncoder(V(), 12, 32, 16, 2, 0.1, bidirectional=False)
    assert bi.out_proj.in_features == 32 and uni.out_proj.in_features == 16
    assert len(bi._gru_weights()) == 16 and len(uni._gru_weights()) == 8
    assert [k for k in uni.state_dict() if "reverse" in k] == []
