"""fp32 mode of the hot path (BASELINE north star: "within 1e-3 fp32"): the autograd Functions of `ops.py` restated on
fp32 values between kernels.

Switch: `ops.set_precision("fp32")`, `with ops.precision_mode("fp32"):`, or GTOS_PRECISION=fp32 in the environment.  The
drop-in modules check it on every forward; parameters, inputs, outputs and state_dict are unchanged (they are fp32 in both
modes).

Every matrix product still runs on the tcgen05 GEMM kernels, on split-bf16 operands with K tripled (see
csrc/precise.cu): x = x_hi + x_lo, A B^T ~= A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T as ONE `gtos_gemm_tn` over
[A_hi | A_lo | A_hi] x [B_hi | B_hi | B_lo].  Roles used throughout: activations and transposed weights are staged with
role 0, gradients and weights with role 1, so every GEMM of a forward or backward pass pairs a role-0 with a role-1 operand,
and the buffer staged for a forward GEMM is reused (viewed as [3 rows, k]) by the weight-gradient GEMM of the backward.

The relation attention (graph_transformer.py:122-159) runs as: K-tripled projection of the relation rows into an fp32
[P, 2D] tensor -> gtos_rel_score_f32 (q/k adds, per-head dots) -> the attention core in its three-pass mode (masks,
softmax, dropout, P.V).  A factorised relation (ops.BankedRelation) is gathered into the dense tensor first: this mode is
the accuracy mode, the bf16 mode is the fast one.  No side streams here - the schedule is the plain dependency order.
"""
import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _lib
from . import ops
from .ops import _attn_desc, _need_cuda, _p, _st, _up8, colsum, dropout_f32, gemm_nn, gemm_tn, new_seed_off, rng_state

BF16 = torch.bfloat16
F32 = torch.float32

# debugging / parity tests: a list here receives the ReLU activation pattern (bool, on the host) of every FFN32Fn forward -
# the sub-gradient an implementation takes at a unit whose pre-activation is within rounding of zero is a CHOICE, and the
# parity tests compare gradients given the same choice (tests/test_gpu_fp32_mode.py, oracle/gtos_oracle.py:_relu)
relu_trace = None


def split3(x2d, role, out=None):
    """fp32 [rows, cols] (any strides) -> the K-tripled bf16 operand [rows, 3 * up8(cols)]: (hi, lo, hi) for role 0,
    (hi, hi, lo) for role 1"""
    _need_cuda(x2d)
    if x2d.dtype != F32:
        raise TypeError(f"split3 expects fp32, got {x2d.dtype}")
    rows, cols = x2d.shape
    kp = _up8(cols)
    if out is None:
        out = torch.empty(rows, 3 * kp, dtype=BF16, device=x2d.device)
    elif tuple(out.shape) != (rows, 3 * kp) or out.dtype != BF16 or not out.is_contiguous():
        raise ValueError("split3: out must be a contiguous bf16 [rows, 3 * up8(cols)] tensor")
    _lib.check(_lib.load().gtos_split3(_p(x2d), x2d.stride(0), x2d.stride(1), rows, cols, _p(out), 3 * kp, kp, role, _st()),
               "split3")
    return out


def _stack(x3):
    """[rows, 3 kp] -> the same memory as [3 rows, kp]: the row-stacked operand of gtos_gemm_nn"""
    rows, k3 = x3.shape
    return x3.view(rows * 3, k3 // 3)


def mm3(x2d, W, b=None, relu=False):
    """x2d @ W^T (+ b) on split operands, no autograd: fp32 [rows, W.shape[0]]"""
    y, _ = gemm_tn(split3(x2d, 0), split3(W.detach(), 1), W.shape[0], bias=b, relu=relu)
    return y


def mm3_acc(A3, B3, out):
    """out[M,N] += A3 @ B3^T in place, through the TMA-store epilogue with `out` as the addend (gtos_gemm_tn_add).  The
    accumulate flag of gtos_gemm_tn takes the per-row read-modify-write epilogue instead: 3x slower on the P-row GEMMs
    (ncu launch list of the first fp32-mode step: 1.18 ms against 0.39 ms for the same product)."""
    M, N = out.shape
    _lib.check(_lib.load().gtos_gemm_tn_add(_p(A3), A3.stride(0), _p(B3), B3.stride(0), None, _p(out), out.stride(0), _p(out),
                                            out.stride(0), M, N, min(A3.shape[1], B3.shape[1]), _st()), "gemm_tn_add")
    return out


def _wgrad(dys, xs, n_out, n_in, out=None):
    """dW [n_out, n_in] = dy^T x from the staged operands (dys role 1, xs role 0)"""
    return gemm_nn(_stack(dys), _stack(xs), n_out, n_in, out=out)


# --------------------------------------------------------------------------------------------
# Linear / FFN
# --------------------------------------------------------------------------------------------
class Linear32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        _need_cuda(x, W)
        shape = x.shape
        K, N = shape[-1], W.shape[0]
        xs = split3(x.contiguous().view(-1, K), 0)
        y, _ = gemm_tn(xs, split3(W.detach(), 1), N, bias=b)
        ctx.save_for_backward(xs, W)
        ctx.meta = (shape, N, K, b is not None)
        return y.view(*shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xs, W = ctx.saved_tensors
        shape, N, K, has_b = ctx.meta
        dy2 = dy.contiguous().view(-1, N)
        dys = split3(dy2, 1)
        dW = _wgrad(dys, xs, N, K)
        db = colsum(dy2) if has_b else None
        dx, _ = gemm_tn(dys, split3(W.detach().t(), 0), K)
        return dx.view(shape), dW, db


class FFN32Fn(torch.autograd.Function):
    """fc2(dropout(relu(fc1 x)))  (graph_transformer.py:60-63, transformer.py:66-69)"""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, p):
        _need_cuda(x, W1, W2)
        shape = x.shape
        D, Fd = shape[-1], W1.shape[0]
        xs = split3(x.contiguous().view(-1, D), 0)
        h, _ = gemm_tn(xs, split3(W1.detach(), 1), Fd, bias=b1, relu=True)
        if relu_trace is not None:
            relu_trace.append((h > 0).view(*shape[:-1], Fd).cpu())
        seed, off = (rng_state(x.device), new_seed_off()) if p > 0 else (None, 0)
        if p > 0:
            dropout_f32(h, p, seed, off, out=h)
        hs = split3(h, 0)
        y, _ = gemm_tn(hs, split3(W2.detach(), 1), D, bias=b2)
        ctx.save_for_backward(xs, h, hs, W1, W2)
        ctx.meta = (p, shape, Fd)
        return y.view(shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xs, h, hs, W1, W2 = ctx.saved_tensors
        p, shape, Fd = ctx.meta
        D = shape[-1]
        dy2 = dy.contiguous().view(-1, D)
        dys = split3(dy2, 1)
        dW2 = _wgrad(dys, hs, D, Fd)
        db2 = colsum(dy2)
        dh, _ = gemm_tn(dys, split3(W2.detach().t(), 0), Fd)
        _lib.check(_lib.load().gtos_relu_drop_bwd_f32(_p(dh), _p(h), _p(dh), dh.numel(), p, _st()), "relu_drop_bwd_f32")
        dhs = split3(dh, 1)
        dW1 = _wgrad(dhs, xs, Fd, D)
        db1 = colsum(dh)
        dx, _ = gemm_tn(dhs, split3(W1.detach().t(), 0), D)
        return dx.view(shape), dW1, db1, dW2, db2, None


# --------------------------------------------------------------------------------------------
# relation-aware self-attention (graph_transformer.py:93-174)
# --------------------------------------------------------------------------------------------
class RelAttn32Fn(torch.autograd.Function):
    """x [N,B,D], relation [N,N,B,D] fp32 (+ rel3: its staged split operand, shared by the layers of one encoder pass);
    returns (out, weights[B,H,N,N] | None)."""

    @staticmethod
    def forward(ctx, x, relation, rel3, key_pad, attn_mask, W_in, b_in, W_rel, W_out, b_out, H, p, need_weights,
                rel_token, rel_acc, weights_dropout):
        _need_cuda(x, relation, W_in)
        lib = _lib.load()
        N, B, D = x.shape
        if tuple(relation.shape) != (N, N, B, D):
            raise ValueError(f"relation must be [N,N,B,D]=[{N},{N},{B},{D}], got {tuple(relation.shape)}")
        hd = D // H
        dev = x.device
        NB, P = N * B, N * N * B
        xs = split3(x.contiguous().view(NB, D), 0)
        qkv, _ = gemm_tn(xs, split3(W_in.detach(), 1), 3 * D, bias=b_in)               # [NB, 3D] fp32: q | k | v
        if rel3 is None:
            rel3 = split3(relation.detach().contiguous().view(P, D), 0)
        PR, _ = gemm_tn(rel3, split3(W_rel.detach(), 1), 2 * D)                        # [P, 2D] fp32: ra | rb   (:122)
        scores = torch.empty(B, H, N, N, dtype=F32, device=dev)                        # [b,h,j,i]
        _lib.check(lib.gtos_rel_score_f32(_p(PR), 2 * D, qkv.data_ptr(), qkv.data_ptr() + 4 * D, 3 * D, _p(scores), N, B, D,
                                          H, _st()), "rel_score_f32")
        probs = torch.empty(B, H, N, N, dtype=F32, device=dev)                         # [b,h,i,j]
        wts = torch.empty(B, H, N, N, dtype=F32, device=dev) if need_weights else None
        att = torch.empty(NB, D, dtype=F32, device=dev)
        seed, off = (rng_state(dev), new_seed_off()) if p > 0 else (None, 0)
        p_w = p if weights_dropout else 0.0
        d = _attn_desc(N, N, B, H, hd)
        d.precise = 1
        d.v, d.ldv = qkv.data_ptr() + 8 * D, 3 * D
        d.scale, d.p_drop = 1.0, p_w
        d.scores_jt, d.key_pad, d.attn_mask = _p(scores), _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs, d.probs_dropped = _p(probs), _p(wts)
        d.out, d.ldo = _p(att), D
        _lib.check(lib.gtos_attn_fwd(C.byref(d), _st()), "attn_fwd(enc, fp32 mode)")
        off2 = 0
        if not weights_dropout and p > 0:                                              # :160-161
            off2 = new_seed_off()
            dropout_f32(att, p, seed, off2, out=att)
        atts = split3(att, 0)
        out, _ = gemm_tn(atts, split3(W_out.detach(), 1), D, bias=b_out)
        ctx.save_for_backward(xs, rel3, qkv, PR, probs, atts, W_in, W_rel, W_out, key_pad, attn_mask)
        ctx.meta = (N, B, D, H, p, seed, off, p_w, off2)
        ctx.rel_acc = rel_acc if (rel_token is not None and rel_token.requires_grad) else None
        ctx.set_materialize_grads(False)
        if wts is None:
            return out.view(N, B, D), None
        return out.view(N, B, D), wts

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dwts):
        xs, rel3, qkv, PR, probs, atts, W_in, W_rel, W_out, key_pad, attn_mask = ctx.saved_tensors
        N, B, D, H, p, seed, off, p_w, off2 = ctx.meta
        lib = _lib.load()
        hd = D // H
        dev = xs.device
        NB, P = N * B, N * N * B
        if dout is None:
            dout = torch.zeros(N, B, D, dtype=F32, device=dev)
        dout2 = dout.contiguous().view(NB, D)
        douts = split3(dout2, 1)
        dW_out = _wgrad(douts, atts, D, D)
        db_out = colsum(dout2)
        datt, _ = gemm_tn(douts, split3(W_out.detach().t(), 0), D)
        if off2:
            dropout_f32(datt, p, seed, off2, out=datt)
        dqkv = torch.empty(NB, 3 * D, dtype=F32, device=dev)
        ds_jt = torch.empty(B, H, N, N, dtype=F32, device=dev)
        ds_ts = torch.empty(B, H, N, N, dtype=F32, device=dev)
        d = _attn_desc(N, N, B, H, hd)
        d.precise = 1
        d.v, d.ldv = qkv.data_ptr() + 8 * D, 3 * D
        d.scale, d.p_drop = 1.0, p_w
        d.key_pad, d.attn_mask = _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs = _p(probs)
        d.dout, d.lddo = _p(datt), D
        d.dprobs_extra = _p(dwts.contiguous()) if dwts is not None else None
        d.dscores_jt, d.dscores_ts = _p(ds_jt), _p(ds_ts)
        d.dv, d.lddv = dqkv.data_ptr() + 8 * D, 3 * D
        _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(enc, fp32 mode)")
        G = torch.empty(P, 2 * D, dtype=F32, device=dev)                               # d ra | d rb per pair
        _lib.check(lib.gtos_rel_grad_f32(_p(PR), 2 * D, qkv.data_ptr(), qkv.data_ptr() + 4 * D, 3 * D, _p(ds_jt), _p(G),
                                         2 * D, N, B, D, H, _st()), "rel_grad_f32")
        _lib.check(lib.gtos_rel_dqk_f32(_p(G), 2 * D, dqkv.data_ptr(), dqkv.data_ptr() + 4 * D, 3 * D, N, B, D, _st()),
                   "rel_dqk_f32")
        Gs = split3(G, 1)
        del G
        d_rel = None
        if ctx.rel_acc is not None or ctx.needs_input_grad[1]:
            Wrt = split3(W_rel.detach().t(), 0)                                        # [D, 3 * 2D]
            if ctx.rel_acc is not None:
                # all layers share one relation tensor: accumulate into the buffer RelTokenFn hands to autograd once
                acc = ctx.rel_acc
                first = acc.buf is None
                if first:
                    acc.buf = torch.empty(N, N, B, D, dtype=F32, device=dev)
                if first:
                    gemm_tn(Gs, Wrt, D, out=acc.buf.view(P, D))
                else:
                    mm3_acc(Gs, Wrt, acc.buf.view(P, D))
            else:
                d_rel, _ = gemm_tn(Gs, Wrt, D)
                d_rel = d_rel.view(N, N, B, D)
        dW_rel = _wgrad(Gs, rel3, 2 * D, D)
        del Gs
        dqkvs = split3(dqkv, 1)
        dW_in = _wgrad(dqkvs, xs, 3 * D, D)
        db_in = colsum(dqkv)
        dx, _ = gemm_tn(dqkvs, split3(W_in.detach().t(), 0), D)
        return (dx.view(N, B, D), d_rel, None, None, None, dW_in, db_in, dW_rel, dW_out, db_out, None, None, None, None,
                None, None)


# --------------------------------------------------------------------------------------------
# vanilla multi-head attention (transformer.py:98-173)
# --------------------------------------------------------------------------------------------
class MHA32Fn(torch.autograd.Function):
    """query [T,B,D]; key [S,B,D] (value is key); returns (out, weights[B,H,T,S] | None)."""

    @staticmethod
    def forward(ctx, query, key, self_attn, key_pad, attn_mask, W_in, b_in, W_out, b_out, H, p, weights_dropout,
                need_weights):
        _need_cuda(query, key, W_in)
        lib = _lib.load()
        T, B, D = query.shape
        S = key.shape[0]
        hd = D // H
        dev = query.device
        qs = split3(query.contiguous().view(T * B, D), 0)
        Wis = split3(W_in.detach(), 1)                                                 # [3D, 3 * up8(D)]
        if self_attn:
            ks = qs
            proj, _ = gemm_tn(qs, Wis, 3 * D, bias=b_in)
            qp, kp, vp, ldq, ldk = proj.data_ptr(), proj.data_ptr() + 4 * D, proj.data_ptr() + 8 * D, 3 * D, 3 * D
            keep = (proj,)
        else:
            ks = split3(key.contiguous().view(S * B, D), 0)
            pq, _ = gemm_tn(qs, Wis[:D], D, bias=b_in[:D])
            pkv, _ = gemm_tn(ks, Wis[D:], 2 * D, bias=b_in[D:])
            qp, kp, vp, ldq, ldk = pq.data_ptr(), pkv.data_ptr(), pkv.data_ptr() + 4 * D, D, 2 * D
            keep = (pq, pkv)
        probs = torch.empty(B, H, T, S, dtype=F32, device=dev)
        p_w = p if weights_dropout else 0.0
        wts = torch.empty(B, H, T, S, dtype=F32, device=dev) if need_weights else None
        att = torch.empty(T * B, D, dtype=F32, device=dev)
        seed, off = (rng_state(dev), new_seed_off()) if p > 0 else (None, 0)
        d = _attn_desc(T, S, B, H, hd)
        d.precise = 1
        d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = qp, ldq, kp, ldk, vp, ldk
        d.scale, d.p_drop = hd ** -0.5, p_w
        d.key_pad, d.attn_mask = _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs, d.probs_dropped = _p(probs), _p(wts)
        d.out, d.ldo = _p(att), D
        _lib.check(lib.gtos_attn_fwd(C.byref(d), _st()), "attn_fwd(dec, fp32 mode)")
        off2 = 0
        if not weights_dropout and p > 0:                                              # transformer.py:156-157
            off2 = new_seed_off()
            dropout_f32(att, p, seed, off2, out=att)
        atts = split3(att, 0)
        out, _ = gemm_tn(atts, split3(W_out.detach(), 1), D, bias=b_out)
        ctx.save_for_backward(qs, ks, probs, atts, W_in, W_out, key_pad, attn_mask, *keep)
        ctx.meta = (T, S, B, D, H, p, p_w, seed, off, off2, self_attn)
        ctx.set_materialize_grads(False)
        if wts is None:
            return out.view(T, B, D), None
        return out.view(T, B, D), wts

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dwts):
        qs, ks, probs, atts, W_in, W_out, key_pad, attn_mask, *keep = ctx.saved_tensors
        T, S, B, D, H, p, p_w, seed, off, off2, self_attn = ctx.meta
        lib = _lib.load()
        hd = D // H
        dev = qs.device
        if dout is None:
            dout = torch.zeros(T, B, D, dtype=F32, device=dev)
        dout2 = dout.contiguous().view(T * B, D)
        douts = split3(dout2, 1)
        dW_out = _wgrad(douts, atts, D, D)
        db_out = colsum(dout2)
        datt, _ = gemm_tn(douts, split3(W_out.detach().t(), 0), D)
        if off2:
            dropout_f32(datt, p, seed, off2, out=datt)
        d = _attn_desc(T, S, B, H, hd)
        d.precise = 1
        if self_attn:
            (proj,) = keep
            dproj = torch.empty(T * B, 3 * D, dtype=F32, device=dev)
            d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = (proj.data_ptr(), 3 * D, proj.data_ptr() + 4 * D, 3 * D,
                                                  proj.data_ptr() + 8 * D, 3 * D)
            d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = (dproj.data_ptr(), 3 * D, dproj.data_ptr() + 4 * D, 3 * D,
                                                        dproj.data_ptr() + 8 * D, 3 * D)
        else:
            pq, pkv = keep
            dpq = torch.empty(T * B, D, dtype=F32, device=dev)
            dpkv = torch.empty(S * B, 2 * D, dtype=F32, device=dev)
            d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = pq.data_ptr(), D, pkv.data_ptr(), 2 * D, pkv.data_ptr() + 4 * D, 2 * D
            d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = (dpq.data_ptr(), D, dpkv.data_ptr(), 2 * D,
                                                        dpkv.data_ptr() + 4 * D, 2 * D)
        ds_ts = torch.empty(B, H, T, S, dtype=F32, device=dev)
        d.scale, d.p_drop = hd ** -0.5, p_w
        d.key_pad, d.attn_mask = _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs = _p(probs)
        d.dout, d.lddo = _p(datt), D
        d.dprobs_extra = _p(dwts.contiguous()) if dwts is not None else None
        d.dscores_ts = _p(ds_ts)
        _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(dec, fp32 mode)")
        dW_in = torch.empty(3 * D, D, dtype=F32, device=dev)
        Wd = W_in.detach()
        if self_attn:
            dprojs = split3(dproj, 1)
            _wgrad(dprojs, qs, 3 * D, D, out=dW_in)
            db_in = colsum(dproj)
            dq_in, _ = gemm_tn(dprojs, split3(Wd.t(), 0), D)
            dk_in = None
        else:
            dpqs, dpkvs = split3(dpq, 1), split3(dpkv, 1)
            _wgrad(dpqs, qs, D, D, out=dW_in[:D])
            _wgrad(dpkvs, ks, 2 * D, D, out=dW_in[D:])
            db_in = torch.cat([colsum(dpq), colsum(dpkv)])
            dq_in, _ = gemm_tn(dpqs, split3(Wd[:D].t(), 0), D)
            dk_in, _ = gemm_tn(dpkvs, split3(Wd[D:].t(), 0), D)
            dk_in = dk_in.view(S, B, D)
        return (dq_in.view(T, B, D), dk_in, None, None, None, dW_in, db_in, dW_out, db_out, None, None, None, None)


# --------------------------------------------------------------------------------------------
# RelationEncoder: embedding -> packed 2-layer bidirectional GRU -> Linear  (encoder.py:90-119)
# --------------------------------------------------------------------------------------------
class GRUBank32Fn(torch.autograd.Function):
    """tokens [Lmax,R] int64 (0 = pad), lengths [R] int64 -> [R, embed_dim]; `weights` = (w_ih, w_hh, b_ih, b_hh) per
    (layer, direction) in nn.GRU order - two directions per layer, or one (bidirectional=False, encoder.py:67).  Per (layer, direction): ONE K-tripled GEMM for the input projections of all time
    steps, then per step a K-tripled GEMM h W_hh^T and the fp32 gate kernel (packed-sequence masking by `lengths`: a finished
    path keeps its state, so the state after the last step is the path's final state in both directions)."""

    @staticmethod
    def forward(ctx, tokens, lengths, embed_w, out_w, out_b, num_layers, hidden, p, *weights):
        _need_cuda(tokens, lengths, embed_w)
        lib = _lib.load()
        dev = embed_w.device
        Lmax, R = tokens.shape
        E = embed_w.shape[1]
        Hh = hidden
        rows = Lmax * R
        ndir = len(weights) // (4 * num_layers)
        if ndir not in (1, 2) or ndir * 4 * num_layers != len(weights):
            raise ValueError(f"GRUBank32Fn: {len(weights)} weight tensors for {num_layers} layers")
        tokens = tokens.contiguous()
        lengths = lengths.contiguous()
        seed = rng_state(dev) if p > 0 else None
        off_e = new_seed_off() if p > 0 else 0
        X = torch.empty(rows, E, dtype=F32, device=dev)
        _lib.check(lib.gtos_embed_gather(_p(embed_w), _p(tokens), rows, E, _p(X), None, E, p, _p(seed), off_e, _st()),
                   "embed_gather")
        saved, layer_offs = [], []
        finals = torch.empty(R, ndir * Hh, dtype=F32, device=dev)
        Kin = E
        for l in range(num_layers):
            last = l == num_layers - 1
            Xs = split3(X, 0)
            out_l = torch.empty(rows, ndir * Hh, dtype=F32, device=dev) if not last else None
            for d in range(ndir):
                w_ih, w_hh, b_ih, b_hh = weights[(l * ndir + d) * 4:(l * ndir + d) * 4 + 4]
                gi, _ = gemm_tn(Xs, split3(w_ih.detach(), 1), 3 * Hh, bias=b_ih)       # [rows, 3H], every time step
                Whs = split3(w_hh.detach(), 1)
                hs = torch.empty(Lmax + 1, R, Hh, dtype=F32, device=dev)               # state before processing step s
                hs[0].zero_()
                hs3 = torch.empty(Lmax, R, 3 * _up8(Hh), dtype=BF16, device=dev)       # its staged operand: this step's GEMM
                gates = torch.empty(Lmax, R, 4 * Hh, dtype=F32, device=dev)            # and the backward's dW_hh GEMM
                for s in range(Lmax):
                    t = s if d == 0 else Lmax - 1 - s
                    gh, _ = gemm_tn(split3(hs[s], 0, out=hs3[s]), Whs, 3 * Hh, bias=b_hh)
                    gi_t = gi[t * R:(t + 1) * R]
                    out_t = out_l[t * R:(t + 1) * R, d * Hh:(d + 1) * Hh] if out_l is not None else None
                    _lib.check(lib.gtos_gru_gate_fwd_f32(_p(gi_t), 3 * Hh, _p(gh), 3 * Hh, _p(hs[s]), _p(lengths), t,
                                                         _p(hs[s + 1]), _p(out_t), ndir * Hh, _p(gates[s]), R, Hh, _st()),
                               "gru_gate_fwd_f32")
                finals[:, d * Hh:(d + 1) * Hh].copy_(hs[Lmax])
                saved += [Xs, gates, hs, hs3, w_ih, w_hh]
            off_l = 0
            if not last and p > 0:                                                     # nn.GRU inter-layer dropout
                off_l = new_seed_off()
                dropout_f32(out_l, p, seed, off_l, out=out_l)
            layer_offs.append(off_l)
            X = out_l
            Kin = ndir * Hh
        fs = split3(finals, 0)
        out, _ = gemm_tn(fs, split3(out_w.detach(), 1), out_w.shape[0], bias=out_b)
        ctx.save_for_backward(tokens, lengths, fs, out_w, *saved)
        ctx.meta = (Lmax, R, E, Hh, num_layers, p, seed, off_e, layer_offs, out_w.shape[0], embed_w.shape[0], ndir)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        tokens, lengths, fs, out_w, *saved = ctx.saved_tensors
        Lmax, R, E, Hh, num_layers, p, seed, off_e, layer_offs, Dout, V, ndir = ctx.meta
        lib = _lib.load()
        dev = dout.device
        rows = Lmax * R
        dout = dout.contiguous()
        douts = split3(dout, 1)
        dW_out = _wgrad(douts, fs, Dout, ndir * Hh)
        db_out = colsum(dout)
        dfinals, _ = gemm_tn(douts, split3(out_w.detach().t(), 0), ndir * Hh)          # [R, ndir H]
        wgrads = [None] * (num_layers * ndir * 4)
        d_layer_out = None                                                             # [rows, 2H] fp32
        d_embed = None
        for l in range(num_layers - 1, -1, -1):
            dx = None
            for d in range(ndir):
                Xs, gates, hs, hs3, w_ih, w_hh = saved[(l * ndir + d) * 6:(l * ndir + d) * 6 + 6]
                Kin = w_ih.shape[1]
                base = (l * ndir + d) * 4
                dgi = torch.empty(rows, 3 * Hh, dtype=F32, device=dev)                 # rows in time order t
                dgh = torch.empty(rows, 3 * Hh, dtype=F32, device=dev)                 # rows in step order s
                dgh3 = torch.empty(rows, 3 * _up8(3 * Hh), dtype=BF16, device=dev)     # staged per step, reused by dW_hh
                Wht = split3(w_hh.detach().t(), 0)                                     # [H, 3 * 3H]
                if l == num_layers - 1:
                    dh = dfinals[:, d * Hh:(d + 1) * Hh].contiguous()
                else:
                    dh = torch.zeros(R, Hh, dtype=F32, device=dev)
                dh_part = torch.empty_like(dh)
                for s in range(Lmax - 1, -1, -1):
                    t = s if d == 0 else Lmax - 1 - s
                    dout_t = d_layer_out[t * R:(t + 1) * R, d * Hh:(d + 1) * Hh] if d_layer_out is not None else None
                    dgi_t, dgh_s = dgi[t * R:(t + 1) * R], dgh[s * R:(s + 1) * R]
                    _lib.check(lib.gtos_gru_gate_bwd_f32(_p(dh), _p(dout_t), ndir * Hh, _p(gates[s]), _p(hs[s]), _p(lengths), t,
                                                         _p(dh_part), _p(dgi_t), 3 * Hh, _p(dgh_s), 3 * Hh, R, Hh, _st()),
                               "gru_gate_bwd_f32")
                    dghs = split3(dgh_s, 1, out=dgh3[s * R:(s + 1) * R])
                    # dh <- dgh @ W_hh + (dh + dout_t) * z
                    _lib.check(lib.gtos_gemm_tn_add(_p(dghs), dghs.stride(0), _p(Wht), Wht.stride(0), None, _p(dh_part), Hh,
                                                    _p(dh), Hh, R, Hh, dghs.shape[1], _st()), "gemm_tn_add")
                dgis = split3(dgi, 1)
                wgrads[base + 0] = _wgrad(dgis, Xs, 3 * Hh, Kin)
                wgrads[base + 1] = _wgrad(dgh3, hs3.view(rows, hs3.shape[-1]), 3 * Hh, Hh)
                wgrads[base + 2] = colsum(dgi)
                wgrads[base + 3] = colsum(dgh)
                Wit = split3(w_ih.detach().t(), 0)                                     # [Kin, 3 * 3H]
                if dx is None:
                    dx, _ = gemm_tn(dgis, Wit, Kin)
                else:
                    mm3_acc(dgis, Wit, dx)
            if l > 0:
                if layer_offs[l - 1]:
                    dropout_f32(dx, p, seed, layer_offs[l - 1], out=dx)
                d_layer_out = dx
            else:
                d_embed = torch.zeros(V, E, dtype=F32, device=dev)
                _lib.check(lib.gtos_embed_scatter_add(_p(dx), _p(tokens), rows, E, _p(d_embed), p, _p(seed), off_e, _st()),
                           "embed_scatter_add")
        return (None, None, d_embed, dW_out, db_out, None, None, None, *wgrads)
