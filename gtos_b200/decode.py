"""Incremental beam decode on B200: K/V caches + device-side beam step.

What the reference does per decode step (generator/generator.py:120-167, generator/search.py:114-168):
  * `mem_dict['graph_state'].index_select(1, indices)` copies the graph memory once per live hypothesis, and every one
    of the 5 attention modules that look at the graph (snt layer external_attn, 3 inference layers' external_attn, the
    alignment layer) re-projects K and V of ALL S x Hyp rows -- 99 % of the step's FLOPs;
  * the token prefix states are concatenated, split per beam, index_selected by the surviving parents and re-projected
    for the whole prefix at every step;
  * top-k, candidate merge and the hypotheses live in Python lists with several device->host syncs per step.

Here:
  * `DecodeEngine.set_memory` projects the graph memory ONCE per source graph for all five modules with one tcgen05
    GEMM ([S*B, D] x [D, 5*2D]) into a bf16 cache; hypotheses read it through `src_index` (no copies);
  * token-side K/V rows are appended to per-layer bf16 caches [Tmax, Hyp, 2D]; a beam reorder only rewrites an int32
    ancestry table (`gtos_beam_ancestry`), never the cached rows;
  * attention with T_q = 1 is `gtos_attn_decode` (one warp per hypothesis x head, fp32 math on bf16 cache rows);
  * the work=True log-prob table is `gtos_token_logprob`; top-k + candidate merge run on the device with fixed shapes
    (`BeamSearchDevice`), so a whole step has no host sync and can be replayed as one CUDA graph per position t.

The numerics contract is the module path's (bf16 operands, fp32 accumulation, 1e-2): tests compare the engine's
log-prob tables with the oracle's plain recomputation, and the merge logic with a restatement of search.py pinned by
golden runs of the reference's own Beam class (tests/golden/make_golden_beam.py).
"""
import ctypes as C

import torch

from . import _lib, ops
from .ops import _need_cuda, _p, _st, _up8


def _gemm(A, W, w_row0, N, bias, M, out32=None, out16=None, ld16=0, relu=False, K=None):
    """C[M,N] = A[M,K] @ W[w_row0:w_row0+N, :K]^T + bias;  A, W bf16 row-major;  out32 fp32 [M,N] and/or out16 (raw
    pointer into a bf16 buffer with row pitch ld16)."""
    K = W.shape[1] if K is None else K
    _lib.check(_lib.load().gtos_gemm_tn(A.data_ptr(), A.stride(0), W.data_ptr() + 2 * w_row0 * W.stride(0), W.stride(0),
                                        _p(bias), _p(out32), out32.stride(0) if out32 is not None else 0,
                                        out16 if out16 is not None else None, ld16, M, N, K, int(relu), 0, _st()),
               "gemm_tn(decode)")


class _AttnW:
    """bf16 copies of one MultiheadAttention's weights, made once (the module path re-casts them every call)."""

    def __init__(self, m):
        self.H, self.D, self.hd = m.num_heads, m.embed_dim, m.head_dim
        self.Win, _ = ops.weight_prep(m.in_proj_weight, want_t=False)            # [3D, D]
        self.bin = m.in_proj_bias.detach().float().contiguous()
        self.Wout, _ = ops.weight_prep(m.out_proj.weight, want_t=False)
        self.bout = m.out_proj.bias.detach().float().contiguous()
        self.bq, self.bkv = self.bin[:self.D], self.bin[self.D:]


class _LayerW:
    def __init__(self, layer):
        self.sa = _AttnW(layer.self_attn)
        self.ea = _AttnW(layer.external_attn)
        self.W1, _ = ops.weight_prep(layer.fc1.weight, want_t=False)
        self.W2, _ = ops.weight_prep(layer.fc2.weight, want_t=False)
        self.b1, self.b2 = layer.fc1.bias.detach(), layer.fc2.bias.detach()
        self.F = layer.fc1.weight.shape[0]
        self.ln_attn = (layer.attn_layer_norm.weight.detach(), layer.attn_layer_norm.bias.detach())
        self.ln_ext = (layer.external_layer_norm.weight.detach(), layer.external_layer_norm.bias.detach())
        self.ln_ff = (layer.ff_layer_norm.weight.detach(), layer.ff_layer_norm.bias.detach())


def attn_decode(q, kv, ld_kv, v_off, row_stride, L, H, hd, slot, slot_ld, key_pad, pad_ld, scale, want_probs=False):
    """q fp32 [Hyp, H*hd]; kv = (tensor, element offset) of the bf16 cache.  Returns (out bf16 [Hyp, H*hd], probs|None)."""
    Hyp = q.shape[0]
    dev = q.device
    out_b = torch.empty(Hyp, H * hd, dtype=torch.bfloat16, device=dev)
    probs = torch.empty(Hyp, H, L, dtype=torch.float32, device=dev) if want_probs else None
    kv_t, kv_off = kv
    _lib.check(_lib.load().gtos_attn_decode(Hyp, L, H, hd, _p(q), q.stride(0), kv_t.data_ptr() + 2 * kv_off, ld_kv, v_off,
                                            row_stride, _p(slot), slot_ld, _p(key_pad), pad_ld, scale, None, 0, _p(out_b),
                                            H * hd, _p(probs), _st()), "attn_decode")
    return out_b, probs


class DecodeEngine:
    """One decode step of `Generator.decode_step` (generator.py:120-167) minus the token embedding front-end:
    snt layer(s) -> DecodeLayer(work=True) -> log-prob table, with all K/V cached.

    snt_encoder: gtos_b200.Transformer (with_external=True); decoder: gtos_b200.DecodeLayer.
    max_hyp: rows per step (live hypotheses, fixed); max_steps: positions the token caches hold."""

    def __init__(self, snt_encoder, decoder, max_hyp, max_steps):
        dev = next(decoder.parameters()).device
        _need_cuda(next(decoder.parameters()))
        if snt_encoder.training or decoder.training:
            raise RuntimeError("DecodeEngine is the inference path: call .eval() on the modules first")
        self.dev = dev
        with torch.no_grad():
            self.layers = [_LayerW(l) for l in snt_encoder.layers] + [_LayerW(l) for l in decoder.inference_core.layers]
            self.n_snt = len(snt_encoder.layers)
            tg = decoder.token_generator
            self.align = _AttnW(tg.alignment_layer)
            self.ln_align = (tg.alignment_layer_norm.weight.detach(), tg.alignment_layer_norm.bias.detach())
            self.Wtransfer, _ = ops.weight_prep(tg.transfer.weight, want_t=False)
            self.btransfer = tg.transfer.bias.detach()
            self.Wgen, _ = ops.weight_prep(tg.generator.weight, want_t=False)
            self.bgen = tg.generator.bias.detach()
            self.Wdiv, self.bdiv = tg.diverter.weight.detach(), tg.diverter.bias.detach()
            self.V = tg.generator.weight.shape[0]
            self.tok_dim = tg.transfer.weight.shape[0]
            D = self.D = self.layers[0].sa.D
            # graph-memory K/V projections of every module that attends to the graph, as ONE weight [n_mem*2D, D]
            mods = [lw.ea for lw in self.layers] + [self.align]
            self.n_mem = len(mods)
            self.Wmem = torch.cat([m.Win[D:3 * D] for m in mods], 0).contiguous()
            self.bmem = torch.cat([m.bkv for m in mods], 0).contiguous()
        self.max_hyp, self.max_steps = int(max_hyp), int(max_steps)
        nl = len(self.layers)
        # token-side K/V caches: position-major [Tmax, Hyp, 2D] bf16 per layer, and the ping-pong ancestry tables
        self.cache = torch.zeros(nl, self.max_steps, self.max_hyp, 2 * D, dtype=torch.bfloat16, device=dev)
        self.anc = [torch.zeros(self.max_steps, self.max_hyp, dtype=torch.int32, device=dev) for _ in range(2)]
        self.mem = None

    # ---- per batch of source graphs ---------------------------------------------------------------------------
    def set_memory(self, graph_state, graph_padding_mask, probe, copy_seq, table_width=None):
        """graph_state [S,B,D], graph_padding_mask [S,B] bool, probe [1,B,D], copy_seq [S,B] int64 (generator.py:98-107).
        table_width: 1 + copy_seq.max() if the caller knows it (avoids the one sync of this call).
        A batch with the shapes of the previous one reuses its buffers in place, so captured graphs stay valid."""
        _need_cuda(graph_state, probe, copy_seq)
        S, B, D = graph_state.shape
        width = self.n_mem * 2 * D
        with torch.no_grad():
            if table_width is None:
                table_width = int(copy_seq.max()) + 1
            W = max(int(table_width), self.V)
            m = self.mem
            if m is None or (m["S"], m["B"], m["W"]) != (S, B, W):
                # new buffers: CUDA graphs captured against the old ones (BeamSearchDevice) must not be replayed -
                # they carry the old pointers and the old S; `mem_epoch` lets their owner notice
                self.mem_epoch = getattr(self, "mem_epoch", 0) + 1
                m = self.mem = dict(S=S, B=B, W=W, ld=width,
                                    kv=torch.empty(S * B, width, dtype=torch.bfloat16, device=self.dev),
                                    pad=torch.zeros(S, B, dtype=torch.uint8, device=self.dev),
                                    probe=torch.empty(B, D, dtype=torch.float32, device=self.dev),
                                    copy_seq=torch.empty(S, B, dtype=torch.int64, device=self.dev))
            gb = ops.cast_bf16(graph_state.contiguous().view(S * B, D))
            _gemm(gb, self.Wmem, 0, width, self.bmem, S * B, out16=m["kv"].data_ptr(), ld16=width)
            if graph_padding_mask is not None:
                m["pad"].copy_(graph_padding_mask)
            else:
                m["pad"].zero_()
            m["probe"].copy_(probe.detach().reshape(B, D))
            m["copy_seq"].copy_(copy_seq)

    # ---- one attention sub-layer with a cache -------------------------------------------------------------------
    def _cross(self, aw, mem_slot, xb, Hyp, src_index, want_probs=False):
        m = self.mem
        D = self.D
        q = torch.empty(Hyp, D, dtype=torch.float32, device=self.dev)
        _gemm(xb, aw.Win, 0, D, aw.bq, Hyp, out32=q)
        att, probs = attn_decode(q, (m["kv"], mem_slot * 2 * D), m["ld"], D, m["B"], m["S"], aw.H, aw.hd, src_index, 0,
                                 m["pad"], m["B"], aw.hd ** -0.5, want_probs)
        a = torch.empty(Hyp, D, dtype=torch.float32, device=self.dev)
        _gemm(att, aw.Wout, 0, D, aw.bout, Hyp, out32=a)
        return a, probs

    def _layer(self, li, x, xb, kv_src_b, t, anc, src_index):
        lw = self.layers[li]
        D, Hyp = self.D, x.shape[0]
        aw = lw.sa
        # self-attention: q of the current row; K/V of the new prefix row appended at position t (slot = row index)
        q = torch.empty(Hyp, D, dtype=torch.float32, device=self.dev)
        _gemm(xb, aw.Win, 0, D, aw.bq, Hyp, out32=q)
        cache = self.cache[li]
        _gemm(kv_src_b, aw.Win, D, 2 * D, aw.bkv, Hyp, out16=cache[t].data_ptr(), ld16=2 * D)
        att, _ = attn_decode(q, (cache, 0), 2 * D, D, self.max_hyp, t + 1, aw.H, aw.hd, anc, self.max_hyp, None, 0,
                             aw.hd ** -0.5)
        a = torch.empty(Hyp, D, dtype=torch.float32, device=self.dev)
        _gemm(att, aw.Wout, 0, D, aw.bout, Hyp, out32=a)
        x, xb = ops.add_layer_norm(a, x, *lw.ln_attn)
        a, _ = self._cross(lw.ea, li, xb, Hyp, src_index)
        x, xb = ops.add_layer_norm(a, x, *lw.ln_ext)
        hb = (torch.empty if lw.F % 8 == 0 else torch.zeros)(Hyp, _up8(lw.F), dtype=torch.bfloat16, device=self.dev)
        _gemm(xb, lw.W1, 0, lw.F, lw.b1, Hyp, out16=hb.data_ptr(), ld16=hb.stride(0), relu=True)
        y = torch.empty(Hyp, D, dtype=torch.float32, device=self.dev)
        _gemm(hb, lw.W2, 0, D, lw.b2, Hyp, out32=y)
        return ops.add_layer_norm(y, x, *lw.ln_ff)

    # ---- one decode step -------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, token_repr, src_index, parent, t, topk=None):
        """token_repr [Hyp, D] fp32 (embedded last token of every live hypothesis, generator.py:131-132);
        src_index int32 [Hyp]: source graph of each hypothesis; parent int32 [Hyp] or None: row of the hypothesis'
        prefix in the PREVIOUS step (search.py:72-76); t: position (0-based).  Returns the log-prob table [Hyp, W], or -
        with topk=k - only (values [Hyp,k], token ids int32 [Hyp,k]) of the k best tokens (generator.py:157), in which
        case the table never exists in HBM."""
        if self.mem is None:
            raise RuntimeError("DecodeEngine.step before set_memory")
        Hyp, D = token_repr.shape
        if Hyp > self.max_hyp or t >= self.max_steps:
            raise ValueError(f"DecodeEngine: Hyp={Hyp} (max {self.max_hyp}), t={t} (max {self.max_steps})")
        lib = _lib.load()
        old, new = self.anc[t % 2], self.anc[(t + 1) % 2]             # ping-pong by position: no host-side state
        _lib.check(lib.gtos_beam_ancestry(_p(old), _p(new), self.max_hyp, _p(parent), t, Hyp, _st()), "beam_ancestry")
        x = token_repr.contiguous()
        xb = ops.cast_bf16(x)
        for li in range(self.n_snt):                    # generator.py:133-142: kv = the prefix of this layer's inputs
            x, xb = self._layer(li, x, xb, xb.view(Hyp, -1), t, new, src_index)
        state_b = xb.view(Hyp, -1)                      # token_state row t (generator.py:143-149)
        outs = self.mem["probe"].index_select(0, src_index.long())     # probe of each hypothesis' graph (decoder.py:82)
        ob = ops.cast_bf16(outs)
        for li in range(self.n_snt, len(self.layers)):  # inference_core: query = probe stream, kv = token states
            outs, ob = self._layer(li, outs, ob, state_b, t, new, src_index)
        return self._token_table(outs, ob.view(Hyp, -1), src_index, topk)

    def _token_table(self, outs, ob, src_index, topk=None):
        """TokenGenerator.forward(work=True), decoder.py:30-59."""
        m = self.mem
        Hyp = outs.shape[0]
        a, align = self._cross(self.align, self.n_mem - 1, ob, Hyp, src_index, want_probs=True)
        outs, ob = ops.add_layer_norm(a, outs, *self.ln_align)
        tok = torch.empty(Hyp, self.tok_dim, dtype=torch.float32, device=self.dev)
        _gemm(ob.view(Hyp, -1), self.Wtransfer, 0, self.tok_dim, self.btransfer, Hyp, out32=tok)
        tok = torch.tanh_(tok)
        gate_logits = torch.addmm(self.bdiv, tok, self.Wdiv.t())
        tokb = ops.cast_bf16(tok)
        logits = torch.empty(Hyp, self.V, dtype=torch.float32, device=self.dev)
        _gemm(tokb, self.Wgen, 0, self.V, self.bgen, Hyp, out32=logits)
        if topk is not None:
            return ops.token_topk(logits, gate_logits, align.view(Hyp, m["S"]), m["copy_seq"], src_index, m["W"], topk)
        return ops.token_logprob(logits, gate_logits, align.view(Hyp, m["S"]), m["copy_seq"], src_index, m["W"])


# ------------------------------------------------------------------------------------------------------------
# device-side beam step (generator/search.py:35-111 Beam, :114-168 search_by_batch) with fixed shapes
# ------------------------------------------------------------------------------------------------------------
class _BeamReadout:
    """host-side read-out shared by the two beam-state implementations (once per batch)"""

    @staticmethod
    def _trace(tok, par, t_last, b, slot):
        seq = []
        for t in range(t_last, -1, -1):
            seq.append(int(tok[t, b, slot]))
            slot = int(par[t, b, slot])
        return seq[::-1]

    def k_best(self, k, alpha):
        """`Beam.get_k_best` (search.py:98-102) for every beam -> list (B) of lists of (token ids without <STR>, score)."""
        tok, par = self.tok.cpu(), self.par.cpu()
        n_done, steps = self.n_done.cpu(), self.steps.cpu()
        d_score, d_step, d_par = self.done_score.cpu(), self.done_step.cpu(), self.done_par.cpu()
        score, live = self.score.cpu(), self.live.cpu()
        out = []
        for b in range(self.B):
            hyps = []
            if int(n_done[b]) > 0:
                for r in range(min(int(n_done[b]), self.K)):
                    t_e = int(d_step[b, r])
                    prefix = self._trace(tok, par, t_e - 1, b, int(d_par[b, r])) if t_e > 0 else []
                    hyps.append((prefix + [self.end_id], float(d_score[b, r])))
            else:                                                              # search.py:99-100
                t_last = int(steps[b]) - 1
                for s in range(self.K):
                    if bool(live[b, s]):
                        hyps.append((self._trace(tok, par, t_last, b, s), float(score[b, s])))
            # score / (1 + len(seq)) ** alpha, where the reference's seq also holds <STR>
            hyps.sort(key=lambda x: x[1] / ((2 + len(x[0])) ** alpha), reverse=True)
            out.append(hyps[:k])
        return out

    def active(self):
        """beams that still submit hypotheses (search.py:93-96: not completed())"""
        return (self.n_done < self.K) & (self.steps < self.Tmax)


class BeamStateFused(_BeamReadout):
    """BeamState whose update is ONE kernel (`gtos_beam_update`, one CTA per beam) fed by the fused top-k of
    `gtos_token_topk`; int32 / uint8 state on the device.  Same semantics as BeamState.update (tests compare them)."""

    def __init__(self, B, K, max_time_step, min_time_step, end_id, unk_id, device):
        if K > 16:
            raise ValueError("BeamStateFused: beam size <= 16")
        self.B, self.K, self.Tmax, self.Tmin = B, K, int(max_time_step), int(min_time_step)
        self.end_id, self.unk_id, self.dev = int(end_id), int(unk_id), device
        f = dict(device=device)
        i = dict(dtype=torch.int32, device=device)
        self.score = torch.empty(B, K, **f)
        self.live = torch.empty(B, K, dtype=torch.uint8, device=device)
        self.n_done, self.steps = torch.empty(B, **i), torch.empty(B, **i)
        self.tok, self.par = torch.empty(self.Tmax, B, K, **i), torch.empty(self.Tmax, B, K, **i)
        self.done_score = torch.empty(B, K, **f)
        self.done_step, self.done_par = torch.empty(B, K, **i), torch.empty(B, K, **i)
        self.reset()

    def reset(self):
        for x in (self.score, self.live, self.n_done, self.steps, self.tok, self.par, self.done_step, self.done_par):
            x.zero_()
        self.live[:, 0] = 1
        self.done_score.fill_(float("-inf"))

    def update(self, t, top_val, top_idx, parent_out, last_tok):
        """top_val f32 / top_idx i32 [B*K, K]; writes parent_out i32 [B*K] and last_tok i64 [B*K] for the next step"""
        _lib.check(_lib.load().gtos_beam_update(self.B, self.K, t, self.Tmin, self.Tmax, self.end_id, self.unk_id,
                                                _p(top_val), _p(top_idx), _p(self.score), _p(self.live), _p(self.n_done),
                                                _p(self.steps), _p(self.tok), _p(self.par), _p(self.done_score),
                                                _p(self.done_step), _p(self.done_par), _p(parent_out), _p(last_tok), _st()),
                   "beam_update")


class BeamState(_BeamReadout):
    """All beams of a batch as fixed-shape tensors (B source graphs x K slots), updated IN PLACE so a step can be
    captured in a CUDA graph.  Mirrors `Beam`: live hypotheses occupy slots 0..n_live-1 in candidate-rank order
    (search.py:66-92), completed ones are kept in a [B, K] table in completion order.  Sequences are stored as
    (token, parent slot) back-pointers per step.  Works on any device (the CPU tests drive it with scripted tables)."""

    def __init__(self, B, K, max_time_step, min_time_step, end_id, unk_id, device):
        self.B, self.K, self.Tmax, self.Tmin = B, K, int(max_time_step), int(min_time_step)
        self.end_id, self.unk_id, self.dev = int(end_id), int(unk_id), device
        f = dict(device=device)
        i = dict(dtype=torch.int64, device=device)
        self.score = torch.empty(B, K, **f)
        self.live = torch.empty(B, K, dtype=torch.bool, device=device)
        self.n_done = torch.empty(B, **i)
        self.steps = torch.empty(B, **i)
        self.tok = torch.empty(self.Tmax, B, K, **i)
        self.par = torch.empty(self.Tmax, B, K, **i)
        self.done_score = torch.empty(B, K, **f)
        self.done_step = torch.empty(B, K, **i)      # step at which END was emitted
        self.done_par = torch.empty(B, K, **i)       # slot of the prefix in the arrangement of the step before
        self._j = torch.arange(K, device=device).unsqueeze(0)
        self.reset()

    def reset(self):
        self.score.zero_()
        self.live.zero_()
        self.live[:, 0] = True                       # one initial hypothesis [STR] per beam (generator.py:108-110)
        self.n_done.zero_()
        self.steps.zero_()
        self.tok.zero_()
        self.par.zero_()
        self.done_score.fill_(float("-inf"))
        self.done_step.zero_()
        self.done_par.zero_()

    def update(self, t, table):
        """One `Beam.update` for every beam (search.py:57-92) from the log-prob table [B*K, W] of step t (0-based).
        Returns (parent slot [B,K], token [B,K]) of the new arrangement."""
        B, K, j = self.B, self.K, self._j
        active = self.active()
        row_live = self.live & active.unsqueeze(1)
        top_s, top_t = torch.topk(table.view(B, K, -1), K, dim=-1)            # generator.py:157 (bsz x k)
        neg = torch.full((), float("-inf"), device=self.dev)
        cand = self.score.unsqueeze(-1) + top_s                               # merge_score, search.py:47-55
        cand = torch.where(top_t == self.unk_id, neg, cand).reshape(B, K * K)
        valid = row_live.unsqueeze(-1).expand(B, K, K).reshape(B, K * K)
        # real candidates first (original order kept), then a stable descending sort by score == list.sort(reverse=True)
        # of search.py:66; candidates of dead rows stay behind every real one, including real ones at -inf (UNK)
        order = torch.sort((~valid).to(torch.int8), dim=1, stable=True)[1]
        key_o = torch.where(valid, cand, neg).gather(1, order)
        srt = torch.sort(key_o, dim=1, descending=True, stable=True)[1]
        pick = order.gather(1, srt)[:, :K]                                    # candidate ids, best first
        n_take = torch.minimum(K - self.n_done, valid.sum(1))                 # search.py:67-68
        n_take = torch.where(active, n_take, torch.zeros_like(n_take))
        taken = j < n_take.unsqueeze(1)
        p_slot = pick // K
        c_score = cand.gather(1, pick)
        c_tok = top_t.reshape(B, K * K).gather(1, pick)
        is_end = (c_tok == self.end_id) & taken
        # len(hyp) - 2 >= min_time_step, with len(seq) = t + 2 once the step-t token is appended (search.py:85-87);
        # an END that comes too early is dropped: neither alive nor completed
        completes = is_end & (t >= self.Tmin)
        stays = taken & ~is_end
        done_rank = self.n_done.unsqueeze(1) + torch.cumsum(completes.to(torch.int64), 1) - 1
        done_rank = torch.where(completes, done_rank, torch.full_like(done_rank, K))       # column K = discard
        zf = torch.zeros(B, 1, device=self.dev)
        zi = torch.zeros(B, 1, dtype=torch.int64, device=self.dev)
        self.done_score.copy_(torch.cat([self.done_score, zf], 1).scatter(1, done_rank, c_score)[:, :K])
        self.done_step.copy_(torch.cat([self.done_step, zi], 1).scatter(1, done_rank, torch.full_like(done_rank, t))[:, :K])
        self.done_par.copy_(torch.cat([self.done_par, zi], 1).scatter(1, done_rank, p_slot)[:, :K])
        self.n_done.add_(completes.sum(1))
        live_rank = torch.cumsum(stays.to(torch.int64), 1) - 1
        live_rank = torch.where(stays, live_rank, torch.full_like(live_rank, K))
        zfk = torch.zeros(B, K + 1, device=self.dev)
        zik = torch.zeros(B, K + 1, dtype=torch.int64, device=self.dev)
        new_score = zfk.scatter(1, live_rank, c_score)[:, :K]
        new_tok = zik.scatter(1, live_rank, c_tok)[:, :K]
        new_par = zik.scatter(1, live_rank, p_slot)[:, :K]
        new_live = j < stays.sum(1).unsqueeze(1)
        a1 = active.unsqueeze(1)                                              # finished beams keep their state
        self.score.copy_(torch.where(a1, new_score, self.score))
        self.live.copy_(torch.where(a1, new_live, self.live))
        self.tok[t].copy_(torch.where(a1, new_tok, torch.zeros_like(new_tok)))
        self.par[t].copy_(torch.where(a1, new_par, j.expand(B, K)))
        self.steps.add_(active.to(torch.int64))
        return self.par[t], self.tok[t]


class BeamSearchDevice:
    """`search_by_batch` (search.py:114-168) on the DecodeEngine: every step runs all B*K rows (dead rows are masked),
    so shapes are static, nothing syncs with the host inside a step, and each position t is replayed as one CUDA graph.

    embed_fn(token_ids int64 [B*K], t) -> [B*K, D] fp32: the token embedding front-end (TokenEncoder + position +
    LayerNorm, generator.py:131-132), which is outside the hot path; it must be capturable (no syncs)
    when use_graphs=True."""

    def __init__(self, engine, beam_size, max_time_step, min_time_step, end_id, unk_id, start_id, embed_fn,
                 use_graphs=False, check_every=8, fused=True):
        self.eng, self.K = engine, int(beam_size)
        self.fused = bool(fused) and self.K <= 16     # fused: gtos_token_topk + gtos_beam_update (2 kernels per step)
        self.Tmax, self.Tmin = int(max_time_step), int(min_time_step)
        self.end_id, self.unk_id, self.start_id = end_id, unk_id, start_id
        self.embed_fn = embed_fn
        self.use_graphs, self.check_every = use_graphs, check_every
        self._graphs, self._pool, self.state = {}, None, None
        self._graph_epoch = getattr(engine, "mem_epoch", 0)

    def _alloc(self, B):
        epoch = getattr(self.eng, "mem_epoch", 0)
        if self._graph_epoch != epoch:
            # DecodeEngine.set_memory replaced its buffers (another S, B or table width): the captured per-position
            # graphs point at the freed buffers and have the old S baked in - drop them (run() falls back to eager
            # steps until capture() is called again)
            self._graphs, self._graph_epoch = {}, epoch
        if self.state is not None and self.state.B == B:
            return
        dev, K = self.eng.dev, self.K
        if B * K > self.eng.max_hyp or self.Tmax > self.eng.max_steps:
            raise ValueError("BeamSearchDevice: engine caches too small for this batch / beam / max_time_step")
        cls = BeamStateFused if self.fused else BeamState
        self.state = cls(B, K, self.Tmax, self.Tmin, self.end_id, self.unk_id, dev)
        self.src_index = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(K)
        self.row_base = (torch.arange(B, device=dev, dtype=torch.int64) * K).unsqueeze(1)
        self.parent = torch.zeros(B * K, dtype=torch.int32, device=dev)
        self.last_tok = torch.empty(B, K, dtype=torch.int64, device=dev)
        self._graphs = {}

    def reset(self):
        self.state.reset()
        self.parent.zero_()
        self.last_tok.fill_(self.start_id)

    def _step(self, t):
        x = self.embed_fn(self.last_tok.view(-1), t)
        if self.fused:
            top_val, top_idx = self.eng.step(x, self.src_index, self.parent if t > 0 else None, t, topk=self.K)
            self.state.update(t, top_val, top_idx, self.parent, self.last_tok)
            return
        table = self.eng.step(x, self.src_index, self.parent if t > 0 else None, t)
        par, tok = self.state.update(t, table)
        self.parent.copy_((par + self.row_base).view(-1))
        self.last_tok.copy_(tok)

    def capture(self, steps=None):
        """one CUDA graph per position t (shapes are static); all graphs share one memory pool"""
        B = self.eng.mem["B"]
        self._alloc(B)
        steps = self.Tmax if steps is None else steps
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.reset()
            for t in range(min(2, steps)):                                     # warm-up: lazy handles, func attributes
                self._step(t)
            side.synchronize()
            for t in range(steps):
                if t in self._graphs:
                    continue
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self._pool, stream=side):
                    self._step(t)
                if self._pool is None:
                    self._pool = g.pool()
                self._graphs[t] = g
        torch.cuda.current_stream().wait_stream(side)
        self.reset()

    def run(self, max_steps=None, early_exit=True):
        B = self.eng.mem["B"]
        self._alloc(B)
        self.reset()
        flag = None
        for t in range(self.Tmax if max_steps is None else max_steps):
            if self.use_graphs and t in self._graphs:
                self._graphs[t].replay()
            else:
                self._step(t)
            if early_exit and (t + 1) % self.check_every == 0:                 # the only host sync, every few steps
                if flag is not None and not bool(flag.item()):
                    break
                flag = self.state.active().any()
        return self.state
