"""Thin Python wrappers over the C ABI and the autograd Functions built from them.

Everything here launches kernels from libgtos_b200.so on the current CUDA stream; PyTorch only
owns the memory (torch.empty) and the autograd graph.  No torch math on the data path.
"""
import ctypes as C
import itertools
import os
import math

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.GtosLibraryError("gtos_b200 ops need CUDA tensors (B200 / sm_100a); there is no CPU path")


def _up8(n):
    return (n + 7) // 8 * 8


def as_u8(mask):
    """bool / uint8 mask -> contiguous uint8 view (no copy for bool)."""
    if mask is None:
        return None
    m = mask.contiguous()
    return m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)


# ---- arithmetic mode -----------------------------------------------------------------------------
# "bf16" (default): bf16 tensor-core operands, fp32 accumulation - the north star's 1e-2 tolerance.
# "fp32": every matrix product on split-bf16 operands (three tensor-core passes, ~2^-17 relative error), fp32 values
# between all kernels - the north star's 1e-3 tolerance against the fp32 reference (gtos_b200/ops32.py, csrc/precise.cu).
# Read on every module forward, so it can be switched between calls; a backward pass uses the mode its forward ran in.
_precision = os.environ.get("GTOS_PRECISION", "bf16")
if _precision not in ("bf16", "fp32"):
    raise ValueError(f"GTOS_PRECISION must be bf16 or fp32, got {_precision!r}")


def precision():
    return _precision


def set_precision(mode):
    global _precision
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    _precision = mode


class precision_mode:
    """`with ops.precision_mode("fp32"): loss = model(batch)` - forward passes inside the block run in that mode"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = _precision
        set_precision(self.mode)
        return self

    def __exit__(self, *exc):
        set_precision(self.prev)
        return False


def fp32_mode():
    return _precision == "fp32"


# ---- second stream for work that is off the critical path of a backward pass -----------------
# The weight-gradient GEMM of a Linear (dW = dY^T X) and its input-gradient GEMM (dX = dY W) are independent and
# each is latency-bound at this path's sizes (M = 2.6-3.8 k rows: 120 CTAs, 8 k-blocks).  `fork()` lets the dW side
# run on a second stream; `join()` before the Function returns makes the caller's stream wait for it, so nothing
# escapes a Function unfinished (under CUDA-graph capture these are plain fork/join edges of the graph).
_side_streams = {}
# Measured on config 2 (bench.py, CUDA-graph step, B200): 9.61 ms with, 9.78 ms without (earlier build: 10.12 vs 10.34).
# Every tcgen05 GEMM CTA owns its SM (~200 KB of shared memory), so two GEMMs interleave rather than overlap; the gain
# comes from the SMs a 84-120-tile GEMM leaves idle.  (Capping the TMA ring at 2 stages so that two CTAs fit one SM made
# the step 0.7 ms SLOWER - the main loop is ingest-latency bound.)  GTOS_SIDE_STREAM=0 disables it; the GPU tests pass
# in both modes.
_side_enabled = os.environ.get("GTOS_SIDE_STREAM", "1") == "1"


# The two directions of a bidirectional GRU layer are independent chains of launches whose tile counts do not fill
# whole waves (config 2: 193 row tiles of 128 on 148 SMs = two waves, the second 30 % full).  Issued on two streams the
# second direction's CTAs take the SMs the first one leaves idle (386 tiles = 2.6 waves per pair of steps instead of
# 4).  GTOS_GRU_STREAMS=0 keeps both directions on the caller's stream.
_gru_streams = os.environ.get("GTOS_GRU_STREAMS", "1") == "1"
_rel_streams = os.environ.get("GTOS_REL_STREAMS", "1") == "1"


class _Fork:
    def __init__(self, enabled=None, which=0):
        self.main = torch.cuda.current_stream()
        self.side = None
        if _side_enabled if enabled is None else enabled:
            key = (self.main.device.index, which)
            if key not in _side_streams:
                _side_streams[key] = torch.cuda.Stream(device=self.main.device)
            self.side = _side_streams[key]
        self._ctx = None
        self.used = False

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(self.main)          # everything queued so far (the operands) is visible
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
            self.used = True
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
            self._ctx = None
        return False

    def join(self):
        if self.used:
            self.main.wait_stream(self.side)
            self.used = False

    def join_param(self, *keep):
        """join for side work whose only products are PARAMETER gradients (weight / bias / LayerNorm gradients).  Inside a
        `with deferred_param_grads():` block the caller's stream does not wait here: the join moves to the end of the
        block, so a weight-gradient GEMM never stalls the dependent chain of the backward pass.  `keep`: the tensors the
        side kernels read - held until the deferred join so the caching allocator cannot hand their memory out."""
        if self.used and _deferred["on"]:
            _deferred["streams"][id(self.side)] = (self.main, self.side)
            _deferred["keep"].extend(keep)
            self.used = False
            return
        self.join()


_deferred = {"on": False, "streams": {}, "keep": []}
_defer_enabled = os.environ.get("GTOS_DEFER_GRADS", "1") == "1"


class deferred_param_grads:
    """`with ops.deferred_param_grads(): loss.backward()` - parameter-gradient work forked to the side stream during the
    backward pass is joined ONCE, when the block ends, instead of at the end of every autograd Function.  Contract: inside
    the block nothing on the caller's stream may read a parameter's .grad - so use it only when the gradients start as
    None (a fresh backward, not an accumulation into existing .grad tensors).  GTOS_DEFER_GRADS=0 makes it a no-op."""

    def __enter__(self):
        self.prev = _deferred["on"]
        _deferred["on"] = _defer_enabled and _side_enabled
        return self

    def __exit__(self, *exc):
        _deferred["on"] = self.prev
        if not self.prev:
            for main, side in _deferred["streams"].values():
                main.wait_stream(side)
            _deferred["streams"].clear()
            _deferred["keep"].clear()
        return False


def join_deferred_now():
    """make the current stream wait for the parameter-gradient work deferred so far (a consumer that reads .grad tensors
    in the middle of a `deferred_param_grads` block - e.g. a gradient bucket that is reduced during the backward pass)"""
    if _deferred["streams"]:
        cur = torch.cuda.current_stream()
        for _main, side in _deferred["streams"].values():
            cur.wait_stream(side)
        _deferred["streams"].clear()
        _deferred["keep"].clear()


def fork(enabled=None, which=0):
    """with fork() as f: <launches on the side stream> ... f.join() before the results are handed on.
    Outputs written inside the block must be allocated BEFORE it (on the caller's stream)."""
    return _Fork(enabled, which)


# ---- dropout RNG: one device-resident 64-bit seed + a per-call-site offset -----------------
_rng_state = {}
_rng_counter = itertools.count(1)


def rng_state(device):
    key = (device.type, device.index)
    if key not in _rng_state:
        _rng_state[key] = torch.tensor([torch.initial_seed() & ((1 << 62) - 1)], dtype=torch.int64, device=device)
    return _rng_state[key]


def reseed(seed, device):
    rng_state(device).fill_(seed & ((1 << 62) - 1))


def advance_rng(device):
    """Bump the device seed (call once per step; capturable in a CUDA graph)."""
    rng_state(device).add_(0x9E3779B97F4A7)


def new_seed_off():
    return (next(_rng_counter) * 0xD1B54A32D192ED03) & _MASK64


# --------------------------------------------------------------------------------------------
# primitive wrappers
# --------------------------------------------------------------------------------------------
def cast_bf16(x2d, ld=None):
    """fp32 [rows, cols] -> bf16 [rows, ld] (zero padded)."""
    _need_cuda(x2d)
    x2d = x2d.contiguous()
    rows, cols = x2d.shape
    ld = _up8(cols) if ld is None else ld
    out = torch.empty(rows, ld, dtype=torch.bfloat16, device=x2d.device)
    _lib.check(_lib.load().gtos_cast_bf16(_p(x2d), cols, _p(out), ld, rows, cols, _st()), "cast_bf16")
    return out


def cast_colsum(x2d):
    """fp32 [rows, cols] -> (bf16 [rows, up8(cols)], column sums fp32 [cols]) in one pass."""
    _need_cuda(x2d)
    x2d = x2d.contiguous()
    rows, cols = x2d.shape
    ld = _up8(cols)
    out = torch.empty(rows, ld, dtype=torch.bfloat16, device=x2d.device)
    sums = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    _lib.check(_lib.load().gtos_cast_colsum(_p(x2d), cols, _p(out), ld, _p(sums), rows, cols, _st()), "cast_colsum")
    return out, sums


class _NoFork:
    def join(self):
        pass

    def join_param(self, *keep):
        pass


# the attention backward kernels (and rel_dqk) write bf16 copies of dq / dk / dv next to the fp32 gradients, so the
# in_proj backward GEMMs start without a cast pass; the bias-gradient column sums move to the side stream
_grad_copies = _side_enabled and os.environ.get("GTOS_GRAD_COPIES", "0") == "1"
_grad_tags = os.environ.get("GTOS_GRAD_TAG", "1") == "1"


def operand_or_cast(x2d, xb):
    """(bf16 operand, fp32 column sums, fork to join) of x2d; xb = the bf16 copy its producer wrote, or None."""
    if xb is None:
        xb, sums = cast_colsum(x2d)
        return xb, sums, _NoFork()
    sums = torch.empty(x2d.shape[1], dtype=torch.float32, device=x2d.device)
    with fork() as f:
        colsum(x2d, out=sums)
    return xb, sums, f


stats = {"grad_operand_tagged": 0, "grad_operand_cast": 0}      # which path incoming gradients took (tests, tools)


def grad_operand(dy, dy2):
    """(bf16 operand copy, fp32 column sums, fork to join) of an incoming gradient dy, viewed as dy2 [rows, cols].
    The LayerNorm backward that produced dy already wrote the bf16 copy (AddLayerNormFn.backward tags its result with
    `_gtos_bf16 = (copy, tensor version)`), so the cast drops out of the critical path and the bias-gradient column sums
    run on the side stream; any other gradient (an autograd accumulation, a user tensor) takes the fused cast_colsum."""
    tag = getattr(dy, "_gtos_bf16", None)
    rows, cols = dy2.shape
    if (_side_enabled and _grad_tags and tag is not None and tag[1] == dy._version and tuple(tag[0].shape) == (rows, cols)
            and cols % 8 == 0 and dy2.data_ptr() == dy.data_ptr()):
        sums = torch.empty(cols, dtype=torch.float32, device=dy2.device)
        with fork() as f:
            colsum(dy2, out=sums)
        stats["grad_operand_tagged"] += 1
        return tag[0], sums, f
    dyb, sums = cast_colsum(dy2)
    stats["grad_operand_cast"] += 1
    return dyb, sums, _NoFork()


def _weight_prep_launch(W, Wb, Wt, rel_heads):
    R, Cc = W.shape
    _lib.check(_lib.load().gtos_weight_prep(_p(W), R, Cc, _p(Wb), _up8(Cc), _p(Wt), _up8(R), rel_heads, _st()),
               "weight_prep")


def weight_prep(W, want_b=True, want_t=True, rel_heads=0, plan=True):
    """W fp32 [R,C] -> (Wb bf16 [R, up8(C)], Wt bf16 [C, up8(R)]).  Inside a WeightPrepPlan.step() block the copies
    may already have been made ahead of time on another stream (plan=False: always cast here, now)."""
    _need_cuda(W)
    Wd = W.detach()
    if plan and _prep_active is not None:
        key = (Wd.data_ptr(), tuple(Wd.shape), rel_heads)
        hit = _prep_active.lookup(key, want_b, want_t)
        if hit is not None:
            return hit
        _prep_active.note(key, W, want_b, want_t, rel_heads)
    Wd = Wd.contiguous()
    R, Cc = Wd.shape
    Wb = torch.empty(R, _up8(Cc), dtype=torch.bfloat16, device=W.device) if want_b else None
    Wt = torch.empty(Cc, _up8(R), dtype=torch.bfloat16, device=W.device) if want_t else None
    _weight_prep_launch(Wd, Wb, Wt, rel_heads)
    return Wb, Wt


_prep_active = None
_prep_ahead = os.environ.get("GTOS_PREP_AHEAD", "1") == "1"


class WeightPrepPlan:
    """The bf16 operand copies of the weights (cast, transpose, relation-weight row permutation: ~60 small launches per
    step at config 2) depend on nothing but the parameters, so they need not sit in the dependent chain of the forward
    pass.  The first `with plan.step():` block records which copies the model asks for; every later block issues all of
    them up front on a third stream - beside the RelationEncoder's GRU steps, which leave room on every SM for these
    4 KB CTAs - and weight_prep() hands them out, joining that stream on first use.  The copies are made from the
    parameters' current values at the start of each block and dropped at its end, so an optimizer step between blocks is
    always seen.  GTOS_PREP_AHEAD=0 turns the plan into a no-op."""

    def __init__(self):
        self.items = None          # [(parameter, want_b, want_t, rel_heads)]
        self._log = None
        self._ready = None
        self._fork = None

    # -- used by weight_prep --
    def lookup(self, key, want_b, want_t):
        if self._ready is None:
            return None
        hit = self._ready.get(key)
        if hit is None or (want_b and hit[0] is None) or (want_t and hit[1] is None):
            return None
        if self._fork is not None:
            self._fork.join()
            self._fork = None
        return (hit[0] if want_b else None), (hit[1] if want_t else None)

    def note(self, key, W, want_b, want_t, rel_heads):
        if self._log is not None:
            prev = self._log.get(key)
            if prev is not None:
                want_b, want_t = want_b or prev[1], want_t or prev[2]
            self._log[key] = (W, want_b, want_t, rel_heads)

    def step(self):
        return _PrepStep(self)


class _PrepStep:
    def __init__(self, plan):
        self.plan = plan

    def __enter__(self):
        global _prep_active
        pl = self.plan
        if not _prep_ahead or _prep_active is not None:
            self.plan = None
            return self
        _prep_active = pl
        if pl.items is None:
            pl._log = {}
            return self
        outs = []
        for W, want_b, want_t, rel_heads in pl.items:                  # allocate on the caller's stream
            Wd = W.detach()
            R, Cc = Wd.shape
            outs.append((Wd, torch.empty(R, _up8(Cc), dtype=torch.bfloat16, device=W.device) if want_b else None,
                         torch.empty(Cc, _up8(R), dtype=torch.bfloat16, device=W.device) if want_t else None, rel_heads))
        pl._ready = {}
        with fork(True, which=1) as f:
            for Wd, Wb, Wt, rel_heads in outs:
                if not Wd.is_contiguous():
                    continue
                _weight_prep_launch(Wd, Wb, Wt, rel_heads)
                pl._ready[(Wd.data_ptr(), tuple(Wd.shape), rel_heads)] = (Wb, Wt)
        pl._fork = f
        return self

    def __exit__(self, *exc):
        global _prep_active
        pl = self.plan
        if pl is None:
            return False
        if pl._log is not None:
            if exc[0] is None:
                pl.items = list(pl._log.values())
            pl._log = None
        if pl._fork is not None:                                       # nothing asked for the copies: still rejoin
            pl._fork.join()
            pl._fork = None
        pl._ready = None
        _prep_active = None
        return False


def gemm_tn(A, B, N, bias=None, f32=True, bf16=False, relu=False, out=None, accumulate=False, K=None, b_off=0,
            a_off=0, M=None):
    """C[M,N] = A[M,K] @ B[N,K]^T.  A, B are bf16 2-D (row stride = .stride(0)); offsets in elements."""
    _need_cuda(A, B)
    M = A.shape[0] if M is None else M
    K = min(A.shape[1] - a_off, B.shape[1] - b_off) if K is None else K
    dev = A.device
    o32 = out if out is not None else (torch.empty(M, N, dtype=torch.float32, device=dev) if f32 else None)
    o16 = torch.empty(M, _up8(N), dtype=torch.bfloat16, device=dev) if bf16 else None
    if o16 is not None and _up8(N) != N:
        o16.zero_()
    _lib.check(_lib.load().gtos_gemm_tn(A.data_ptr() + 2 * a_off, A.stride(0), B.data_ptr() + 2 * b_off, B.stride(0),
                                        _p(bias), _p(o32), o32.stride(0) if o32 is not None else 0, _p(o16),
                                        o16.stride(0) if o16 is not None else 0, M, N, K, int(relu), int(accumulate),
                                        _st()), "gemm_tn")
    return o32, o16


def gemm_nn(A, B, M, N, out=None, a_off=0, b_off=0):
    """C[M,N] = sum_k A[k, m] B[k, n];  A [Kd, >=M], B [Kd, >=N] bf16."""
    _need_cuda(A, B)
    Kd = A.shape[0]
    lib = _lib.load()
    ws_elems = lib.gtos_gemm_nn_workspace(M, N, Kd)
    ws = None
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    _lib.check(lib.gtos_gemm_nn(A.data_ptr() + 2 * a_off, A.stride(0), B.data_ptr() + 2 * b_off, B.stride(0), _p(out),
                                out.stride(0), M, N, Kd, _p(ws), ws_elems, _st()), "gemm_nn")
    return out


def zero_regions(tensors):
    """zero several (contiguous) tensors / row-range views with ONE kernel launch instead of one fill each"""
    ts = [t for t in tensors if t is not None and t.numel() > 0]
    if not ts:
        return
    _need_cuda(*ts)
    for t in ts:
        if not t.is_contiguous():
            raise ValueError("zero_regions: contiguous views only")
    n = len(ts)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    sizes = (C.c_int64 * n)(*[t.numel() * t.element_size() for t in ts])
    _lib.check(_lib.load().gtos_zero_regions(ptrs, sizes, n, _st()), "zero_regions")


def colsum(x2d, out=None):
    if out is None:
        out = torch.empty(x2d.shape[1], dtype=torch.float32, device=x2d.device)
    fn = _lib.load().gtos_colsum_bf16 if x2d.dtype == torch.bfloat16 else _lib.load().gtos_colsum
    _lib.check(fn(_p(x2d), x2d.stride(0), _p(out), x2d.shape[0], x2d.shape[1], _st()), "colsum")
    return out


def rel_tiling(N, B, D, H):
    out = (C.c_int32 * 5)()
    _lib.check(_lib.load().gtos_rel_tiling(N, B, D, H, out), "rel_tiling")
    return dict(bi=out[0], bj=out[1], ni_blk=out[2], nj_blk=out[3], tiles=out[4])


def relation_to_bf16(relation):
    """dense relation fp32 [N,N,B,D] -> bf16 copy, made once per encoder pass and shared by all layers."""
    N1, N2, B, D = relation.shape
    return cast_bf16(relation.reshape(N1 * N2 * B, D), ld=D).view(N1, N2, B, D)


def dropout_f32(x, p, seed_ptr, seed_off, out=None):
    x = x.contiguous()
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().gtos_dropout_f32(_p(x), _p(out), x.numel(), p, _p(seed_ptr), seed_off, _st()), "dropout")
    return out


def _attn_desc(T, S, B, H, hd):
    d = _lib.AttnDesc()
    d.T, d.S, d.B, d.H, d.hd = T, S, B, H, hd
    return d


# --------------------------------------------------------------------------------------------
# residual + dropout + LayerNorm
# --------------------------------------------------------------------------------------------
class AddLayerNormFn(torch.autograd.Function):
    """y = LayerNorm(res + dropout(x)); returns (y fp32, y bf16 [non-differentiable operand copy])."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, p):
        _need_cuda(x, res, gamma)
        shape = x.shape
        D = shape[-1]
        x2 = x.contiguous().view(-1, D)
        r2 = res.contiguous().view(-1, D) if res is not None else None
        rows = x2.shape[0]
        dev = x.device
        y = torch.empty(rows, D, dtype=torch.float32, device=dev)
        yb = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
        z = torch.empty(rows, D, dtype=torch.float32, device=dev)
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        seed, off = (rng_state(dev), new_seed_off()) if p > 0 else (None, 0)
        _lib.check(_lib.load().gtos_add_ln_fwd(_p(x2), _p(r2), _p(gamma), _p(beta), _p(y), _p(yb), _p(z), _p(mean),
                                               _p(rstd), rows, D, p, _p(seed), off, _st()), "add_ln_fwd")
        ctx.save_for_backward(z, mean, rstd, gamma)
        ctx.meta = (p, seed, off, shape, res is not None)
        yb = yb.view(shape)
        ctx.mark_non_differentiable(yb)
        ctx.set_materialize_grads(False)            # no zero-filled "gradient" for the bf16 operand copy
        return y.view(shape), yb

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dyb):
        z, mean, rstd, gamma = ctx.saved_tensors
        p, seed, off, shape, has_res = ctx.meta
        rows, D = z.shape
        if dy is None:
            dy = torch.zeros(shape, dtype=torch.float32, device=z.device)
        dy2 = dy.contiguous().view(rows, D)
        dres = torch.empty_like(z)
        dx = torch.empty_like(z) if p > 0 else None
        dgamma = torch.empty(D, dtype=torch.float32, device=z.device)
        dbeta = torch.empty(D, dtype=torch.float32, device=z.device)
        with fork() as f_par:                       # dgamma / dbeta: nothing downstream in this backward pass reads them
            _lib.check(_lib.load().gtos_ln_param_grad(_p(dy2), _p(z), _p(mean), _p(rstd), _p(dgamma), _p(dbeta), rows, D,
                                                      _st()), "ln_param_grad")
        dxb = torch.empty(rows, D, dtype=torch.bfloat16, device=z.device) if D % 8 == 0 else None
        _lib.check(_lib.load().gtos_add_ln_bwd(_p(dy2), _p(z), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dxb),
                                               None, None, rows, D, p, _p(seed), off, _st()), "add_ln_bwd")
        f_par.join_param(dy2, z, mean, rstd)
        dres = dres.view(shape)
        dxo = dres if dx is None else dx.view(shape)
        if dxb is not None:
            dxo._gtos_bf16 = (dxb, dxo._version)    # operand copy for the sublayer's backward GEMMs (ops.grad_operand)
        return dxo, (dres if has_res else None), dgamma, dbeta, None


def add_layer_norm(x, res, gamma, beta, p=0.0):
    return AddLayerNormFn.apply(x, res, gamma, beta, float(p))


# --------------------------------------------------------------------------------------------
# position-wise feed-forward: fc2(dropout(relu(fc1 x)))
# --------------------------------------------------------------------------------------------
class FFNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, xb, W1, b1, W2, b2, p):
        _need_cuda(x, W1, W2)
        shape = x.shape
        D = shape[-1]
        Fd = W1.shape[0]
        xb2 = (xb if xb is not None else cast_bf16(x.contiguous().view(-1, D))).view(-1, _up8(D))
        W1b, W1t = weight_prep(W1)
        W2b, W2t = weight_prep(W2)
        _, hb = gemm_tn(xb2, W1b, Fd, bias=b1, f32=False, bf16=True, relu=True)
        seed, off = (rng_state(x.device), new_seed_off()) if p > 0 else (None, 0)
        if p > 0:
            _lib.check(_lib.load().gtos_dropout_bf16(_p(hb), hb.numel(), p, _p(seed), off, _st()), "dropout_bf16")
        y, _ = gemm_tn(hb, W2b, D, bias=b2)
        ctx.save_for_backward(xb2, hb, W1t, W2t)
        ctx.meta = (p, shape, Fd)
        return y.view(shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xb2, hb, W1t, W2t = ctx.saved_tensors
        p, shape, Fd = ctx.meta
        D = shape[-1]
        dy2 = dy.contiguous().view(-1, D)
        M = dy2.shape[0]
        dyb, db2, f_db = grad_operand(dy, dy2)
        dev = dy.device
        dW2 = torch.empty(D, Fd, dtype=torch.float32, device=dev)
        dW1 = torch.empty(Fd, D, dtype=torch.float32, device=dev)
        db1f = torch.empty(_up8(Fd), dtype=torch.float32, device=dev)
        with fork() as f2:
            gemm_nn(dyb, hb, D, Fd, out=dW2)
        dh, _ = gemm_tn(dyb, W2t, Fd)                       # [M,F] = dy @ W2
        dhb = torch.empty(M, _up8(Fd), dtype=torch.bfloat16, device=dev)
        _lib.check(_lib.load().gtos_relu_drop_bwd(_p(dh), _p(hb), None, _p(dhb), dh.numel(), p, _st()), "relu_drop_bwd")
        with fork() as f1:
            gemm_nn(dhb, xb2, Fd, D, out=dW1)
            colsum(dhb, out=db1f)
        dx, _ = gemm_tn(dhb, W1t, D)
        f2.join_param(dyb, hb)
        f1.join_param(dhb, xb2)
        f_db.join()                                 # reads the incoming gradient itself (autograd may accumulate into that tensor later)
        return dx.view(shape), None, dW1, db1f[:Fd], dW2, db2, None


def ffn(x, xb, W1, b1, W2, b2, p=0.0):
    if _precision == "fp32":
        from . import ops32
        return ops32.FFN32Fn.apply(x, W1, b1, W2, b2, float(p))
    return FFNFn.apply(x, xb, W1, b1, W2, b2, float(p))


# --------------------------------------------------------------------------------------------
# relation-aware multi-head self-attention (graph_transformer.py:93-174)
# --------------------------------------------------------------------------------------------
class RelAttnFn(torch.autograd.Function):
    """inputs x [N,B,D], relation [N,N,B,D] (+ its shared bf16 copy relb); returns (out, weights[B,H,N,N] | None)."""

    @staticmethod
    def forward(ctx, x, xb, relation, relb, key_pad, attn_mask, W_in, b_in, W_rel, W_out, b_out, H, p, need_weights,
                rel_token=None, rel_acc=None, weights_dropout=True, banked=None):
        _need_cuda(x, W_in)
        lib = _lib.load()
        N, B, D = x.shape
        fused_bank = banked is not None and banked_fwd_supported(banked, N, B, D, H)
        if not fused_bank:
            if banked is not None and relb is None:
                relb = banked.relb
            rshape = tuple((relation if relation is not None else relb).shape)
            if rshape != (N, N, B, D):
                raise ValueError(f"relation must be [N,N,B,D]=[{N},{N},{B},{D}], got {rshape}")
        hd = D // H
        dev = x.device
        NB = N * B
        xb2 = (xb if xb is not None else cast_bf16(x.contiguous().view(NB, D))).view(NB, D)
        if relb is None and not fused_bank:
            relb = relation_to_bf16(relation.detach().contiguous())
        Wib, Wit = weight_prep(W_in)
        Wperm, WpermT = weight_prep(W_rel, rel_heads=H)
        Wob, Wot = weight_prep(W_out)
        # One kernel per layer on the dense relation tensor (the north star): projection GEMM + scores + mask + softmax +
        # dropout + P.V (gtos_rel_attn_fwd) when a 128-pair tile holds all keys of its queries; otherwise the score
        # kernel writes [B,H,N,N] scores and the attention core finishes (gtos_rel_score + gtos_attn_fwd).
        fused_dense = (not fused_bank and _rel_fused_fwd and attn_mask is None
                       and lib.gtos_rel_attn_fusable(N, B, D, H) == 1)
        need_grad = any(ctx.needs_input_grad)
        vproj = None
        f_v = _NoFork()
        if fused_dense:
            # q | k | v in ONE bf16 GEMM; the fp32 copy of v that the backward's attention core reads is made beside it
            _, qkb = gemm_tn(xb2, Wib, 3 * D, bias=b_in, f32=False, bf16=True)      # [NB, 3D] bf16
            ldqk = 3 * D
            if need_grad:
                vproj = torch.empty(NB, D, dtype=torch.float32, device=dev)
                with fork() as f_v:
                    gemm_tn(xb2, Wib, D, bias=b_in[2 * D:], K=D, b_off=2 * D * Wib.stride(0), out=vproj)
        else:
            # q,k feed only the fused relation kernels (staged there by TMA as bf16); v feeds the attention core
            _, qkb = gemm_tn(xb2, Wib, 2 * D, bias=b_in[:2 * D], f32=False, bf16=True)  # [NB, 2D] bf16
            ldqk = 2 * D
            vproj = torch.empty(NB, D, dtype=torch.float32, device=dev)
            with fork() as f_v:                     # v is first read by the attention core, after the score kernel
                gemm_tn(xb2, Wib, D, bias=b_in[2 * D:], K=D, b_off=2 * D * Wib.stride(0), out=vproj)   # [NB, D] fp32
        probs = torch.empty(B, H, N, N, dtype=torch.float32, device=dev)           # [b,h,i,j]
        wts = torch.empty(B, H, N, N, dtype=torch.float32, device=dev) if need_weights else None
        att = torch.empty(NB, D, dtype=torch.float32, device=dev)
        attb = torch.empty(NB, D, dtype=torch.bfloat16, device=dev)
        seed, off = (rng_state(dev), new_seed_off()) if p > 0 else (None, 0)
        p_w = p if weights_dropout else 0.0         # graph_transformer.py:154-161: dropout on the weights OR on the output
        PB = None
        if fused_bank:
            # SURVEY 8 f-0, forward half: [ra | rb] of a pair is a ROW of the projected bank - one R-row GEMM, then ONE
            # kernel for gather + scores + masks + softmax + dropout + PV (graph_transformer.py:122-159)
            _, PB = gemm_tn(banked.bankb, Wperm, 2 * D, f32=False, bf16=True)       # [R, 2D] bf16
            f_v.join()
            _lib.check(lib.gtos_rel_attn_banked_fwd(_p(PB), PB.stride(0), _p(banked.idx), qkb.data_ptr(),
                                                    qkb.data_ptr() + 2 * D, ldqk, _p(vproj), D, _p(key_pad), _p(attn_mask),
                                                    p_w, _p(seed), off, _p(probs), _p(wts), _p(att), D, _p(attb), N, B, D, H,
                                                    banked.bankb.shape[0], _st()), "rel_attn_banked_fwd")
            relb = PB                               # saved in relb's slot for the backward
        elif fused_dense:
            _lib.check(lib.gtos_rel_attn_fwd(_p(relb), _p(Wperm), qkb.data_ptr(), qkb.data_ptr() + 2 * D, ldqk,
                                             qkb.data_ptr() + 4 * D, ldqk, _p(key_pad), p_w, _p(seed), off, _p(probs), _p(wts),
                                             _p(att), D, _p(attb), N, B, D, H, _st()), "rel_attn_fwd")
            f_v.join()
        else:
            scores = torch.empty(B, H, N, N, dtype=torch.float32, device=dev)      # [b,h,j,i]
            _lib.check(lib.gtos_rel_score(_p(relb), _p(Wperm), qkb.data_ptr(), qkb.data_ptr() + 2 * D, ldqk, _p(scores),
                                          N, B, D, H, _st()), "rel_score")
            f_v.join()
            d = _attn_desc(N, N, B, H, hd)
            d.v, d.ldv = vproj.data_ptr(), D
            d.scale, d.p_drop = 1.0, p_w
            d.scores_jt, d.key_pad, d.attn_mask = _p(scores), _p(key_pad), _p(attn_mask)
            d.seed_ptr, d.seed_off = _p(seed), off
            d.probs, d.probs_dropped = _p(probs), _p(wts)
            d.out, d.ldo, d.out_bf16 = _p(att), D, _p(attb)
            _lib.check(lib.gtos_attn_fwd(C.byref(d), _st()), "attn_fwd(enc)")
        off2 = 0
        if not weights_dropout and p > 0:
            off2 = new_seed_off()
            dropout_f32(att, p, seed, off2, out=att)
            attb = cast_bf16(att)
        out, _ = gemm_tn(attb, Wob, D, bias=b_out)
        ctx.save_for_backward(xb2, relb, qkb, vproj, probs, attb, Wperm, WpermT, Wit, Wot, key_pad, attn_mask)
        ctx.meta = (N, B, D, H, p, seed, off, p_w, off2, ldqk)
        ctx.fused_bank = banked if fused_bank else None
        ctx.rel_acc = rel_acc if (rel_token is not None and rel_token.requires_grad) else None
        ctx.set_materialize_grads(False)            # unused attention weights: no zero-filled [B,H,N,N] gradient
        if wts is None:
            return out.view(N, B, D), None
        return out.view(N, B, D), wts

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dwts):
        xb2, relb, qkb, vproj, probs, attb, Wperm, WpermT, Wit, Wot, key_pad, attn_mask = ctx.saved_tensors
        N, B, D, H, p, seed, off, p_w, off2, ldqk = ctx.meta
        lib = _lib.load()
        hd = D // H
        if dout is None:
            dout = torch.zeros(N, B, D, dtype=torch.float32, device=xb2.device)
        dev = dout.device
        NB = N * B
        dout2 = dout.contiguous().view(NB, D)
        doutb, db_out, f_db = grad_operand(dout, dout2)
        dW_out = torch.empty(D, D, dtype=torch.float32, device=dev)
        with fork() as f_out:
            gemm_nn(doutb, attb, D, D, out=dW_out)
        datt, _ = gemm_tn(doutb, Wot, D)                                           # [NB, D]
        if off2:
            dropout_f32(datt, p, seed, off2, out=datt)
        dqkv = torch.empty(NB, 3 * D, dtype=torch.float32, device=dev)
        dqkv_b = torch.empty(NB, 3 * D, dtype=torch.bfloat16, device=dev) if (_grad_copies and D % 8 == 0) else None
        dqb_p = dqkv_b.data_ptr() if dqkv_b is not None else None
        dkb_p = dqkv_b.data_ptr() + 2 * D if dqkv_b is not None else None
        ds_jt = torch.empty(B, H, N, N, dtype=torch.float32, device=dev)
        ds_ts = torch.empty(B, H, N, N, dtype=torch.float32, device=dev)
        d = _attn_desc(N, N, B, H, hd)
        d.v, d.ldv = vproj.data_ptr(), D
        d.scale, d.p_drop = 1.0, p_w
        d.key_pad, d.attn_mask = _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs = _p(probs)
        d.dout, d.lddo = _p(datt), D
        d.dprobs_extra = _p(dwts.contiguous()) if dwts is not None else None
        d.dscores_jt, d.dscores_ts = _p(ds_jt), _p(ds_ts)
        d.dv, d.lddv = dqkv.data_ptr() + 8 * D, 3 * D
        d.dv_bf16 = dqkv_b.data_ptr() + 4 * D if dqkv_b is not None else None
        # dV = Pd^T dO needs nothing from the query-side kernel and nothing reads it before the in_proj backward GEMMs:
        # it runs beside the query side and the relation gradient kernels
        with fork(which=4) as f_dv:
            d.bwd_part = 2
            _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(enc, key side)")
        d.bwd_part = 1
        _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(enc, query side)")
        tiles = rel_tiling(N, B, D, H)["tiles"]
        G = torch.empty(tiles * 128, 2 * D, dtype=torch.bfloat16, device=dev)
        if ctx.fused_bank is not None:
            bk = ctx.fused_bank                     # relb's slot holds the projected bank PB [R, 2D]
            _lib.check(lib.gtos_rel_grad_banked(_p(relb), relb.stride(0), _p(bk.idx), qkb.data_ptr(), qkb.data_ptr() + 2 * D,
                                                ldqk, _p(ds_jt), _p(G), N, B, D, H, bk.bankb.shape[0], _st()),
                       "rel_grad_banked")
        else:
            _lib.check(lib.gtos_rel_grad(_p(relb), _p(Wperm), qkb.data_ptr(), qkb.data_ptr() + 2 * D, ldqk, _p(ds_jt),
                                         _p(G), N, B, D, H, _st()), "rel_grad")
        d_rel = None
        if ctx.rel_acc is not None and ctx.rel_acc.banked is not None:
            # relation = bank[idx]: segmented sum of G over bank rows, then R-row GEMMs instead of
            # P-row GEMMs; no [N,N,B,D] gradient tensor, no atomic scatter
            acc = ctx.rel_acc
            bk = acc.banked
            R = bk.bankb.shape[0]
            L = acc.n_layers
            if acc.S is None:
                # the layers' segment sums sit side by side: d bank = [S_1|..|S_L] [Wperm_1;..;Wperm_L] is one GEMM
                acc.S = torch.zeros(R, L * 2 * D, dtype=torch.bfloat16, device=dev)  # rows without pairs stay zero
                acc.Wcat = torch.zeros(D, L * 2 * D, dtype=torch.bfloat16, device=dev)
                acc.spill = torch.empty(R, 2 * D, dtype=torch.float32, device=dev)
                acc.slot = 0
            if acc.slot >= L:
                raise RuntimeError("RelGradAcc: more relation-attention backward passes than layers")
            c0 = acc.slot * 2 * D
            acc.slot += 1
            s_ptr = acc.S.data_ptr() + 2 * c0
            dW_rel = torch.empty(2 * D, D, dtype=torch.float32, device=dev)
            dW_in = torch.empty(3 * D, D, dtype=torch.float32, device=dev)
            # The bank side (segmented sum -> dW_rel, both off the critical path) and the node side (dq/dk segment sums ->
            # in_proj backward -> dx) both stream G from HBM and are independent: two streams (GTOS_REL_STREAMS=0: one).
            with fork(_rel_streams or None, which=2) as f_rel:
                _lib.check(lib.gtos_rel_segsum(_p(G), _p(bk.order), _p(bk.keys), N * N * B, 2 * D, s_ptr, L * 2 * D,
                                               _p(acc.spill), _st()), "rel_segsum")
                _lib.check(lib.gtos_rel_dw_bank(s_ptr, L * 2 * D, _p(bk.bankb), _p(dW_rel), R, D, H, _st()), "rel_dw_bank")
                acc.Wcat[:, c0:c0 + 2 * D].copy_(WpermT[:, :2 * D])
            _lib.check(lib.gtos_rel_dqk(_p(G), dqkv.data_ptr(), dqkv.data_ptr() + 4 * D, 3 * D, dqb_p, dkb_p, N, B, D, H,
                                        _st()), "rel_dqk")
            f_dv.join()
            dqkvb, db_in, f_dbin = operand_or_cast(dqkv, dqkv_b)
            with fork() as f_in:
                gemm_nn(dqkvb, xb2, 3 * D, D, out=dW_in)
            dx, _ = gemm_tn(dqkvb, Wit, D)
            f_out.join_param(doutb, attb)
            f_rel.join()                        # the segment sums feed BankTokenFn's d-bank GEMM on the caller's stream
            f_in.join_param(dqkvb, xb2)
            f_db.join()
            f_dbin.join_param(dqkv)
            return (dx.view(N, B, D), None, None, None, None, None, dW_in, db_in, dW_rel, dW_out, db_out, None, None,
                    None, None, None, None, None)
        _lib.check(lib.gtos_rel_dqk(_p(G), dqkv.data_ptr(), dqkv.data_ptr() + 4 * D, 3 * D, dqb_p, dkb_p, N, B, D, H, _st()),
                   "rel_dqk")
        if ctx.rel_acc is not None:
            # all layers share one relation tensor: reduce this layer's gradient straight into the shared buffer
            # (TMA reduce-add), which RelTokenFn hands to autograd once
            acc = ctx.rel_acc
            first = acc.buf is None
            if first:
                acc.buf = torch.empty(N, N, B, D, dtype=torch.float32, device=dev)
            _lib.check(lib.gtos_rel_drel(_p(G), _p(WpermT), _p(acc.buf), 0 if first else 1, N, B, D, H, _st()), "rel_drel")
        elif ctx.needs_input_grad[2]:
            d_rel = torch.empty(N, N, B, D, dtype=torch.float32, device=dev)
            _lib.check(lib.gtos_rel_drel(_p(G), _p(WpermT), _p(d_rel), 0, N, B, D, H, _st()), "rel_drel")
        ws_elems = lib.gtos_rel_dw_workspace(N, B, D, H)
        ws = None
        dW_rel = torch.empty(2 * D, D, dtype=torch.float32, device=dev)
        _lib.check(lib.gtos_rel_dw(_p(G), _p(relb), _p(dW_rel), _p(ws), ws_elems, N, B, D, H, _st()), "rel_dw")
        f_dv.join()
        dqkvb, db_in, f_dbin = operand_or_cast(dqkv, dqkv_b)
        dW_in = torch.empty(3 * D, D, dtype=torch.float32, device=dev)
        with fork() as f_in:
            gemm_nn(dqkvb, xb2, 3 * D, D, out=dW_in)
        dx, _ = gemm_tn(dqkvb, Wit, D)
        f_out.join_param(doutb, attb)
        f_in.join_param(dqkvb, xb2)
        f_db.join()
        f_dbin.join_param(dqkv)
        return (dx.view(N, B, D), None, d_rel, None, None, None, dW_in, db_in, dW_rel, dW_out, db_out, None, None,
                None, None, None, None, None)


def rel_attention_composed(query, key, value, relation, key_padding_mask, attn_mask, W_in, b_in, W_rel, W_out, b_out, H, p,
                           weights_dropout, need_weights):
    """RelationMultiheadAttention.forward for the argument combinations gtos itself never uses (key / value that are
    not the query tensor): the five projections run on the tcgen05 GEMM (LinearFn), the small remainder is written with
    PyTorch CUDA ops exactly as graph_transformer.py:118-172 states it.  CUDA only."""
    import torch.nn.functional as F
    T, B, D = query.shape
    S = key.shape[0]
    hd = D // H
    q = linear(query, W_in[:D], b_in[:D])
    k = linear(key, W_in[D:2 * D], b_in[D:2 * D])
    v = linear(value, W_in[2 * D:], b_in[2 * D:])
    q = q.contiguous().view(T, B * H, hd)
    k = k.contiguous().view(S, B * H, hd)
    v = v.contiguous().view(S, B * H, hd)
    ra, rb = linear(relation, W_rel).chunk(2, dim=-1)
    ra = ra.contiguous().view(T, S, B * H, hd).transpose(0, 1)
    rb = rb.contiguous().view(T, S, B * H, hd).transpose(0, 1)
    q = (q.unsqueeze(1) + ra) * (hd ** -0.5)
    k = k.unsqueeze(0) + rb
    w = torch.einsum('ijbn,ijbn->ijb', q, k)
    if attn_mask is not None:
        w = w.masked_fill(attn_mask.bool().unsqueeze(-1), float('-inf'))
    if key_padding_mask is not None:
        w = w.view(T, S, B, H).masked_fill(key_padding_mask.bool().unsqueeze(0).unsqueeze(-1), float('-inf')).view(T, S, B * H)
    w = F.softmax(w, dim=1)
    if weights_dropout:
        w = F.dropout(w, p=p, training=p > 0)
    attn = torch.einsum('ijb,jbn->bin', w, v)
    if not weights_dropout:
        attn = F.dropout(attn, p=p, training=p > 0)
    attn = attn.transpose(0, 1).contiguous().view(T, B, D)
    out = linear(attn, W_out, b_out)
    return out, (w.view(T, S, B, H) if need_weights else None)


def mha_composed(query, key, value, key_padding_mask, attn_mask, W_in, b_in, W_out, b_out, H, p, weights_dropout,
                 need_weights):
    """MultiheadAttention.forward with `key is not value` (transformer.py:113-118, never used by gtos): projections on
    the tcgen05 GEMM, the rest as transformer.py:119-171 states it, in PyTorch CUDA ops.  CUDA only."""
    import torch.nn.functional as F
    T, B, D = query.shape
    hd = D // H
    q = linear(query, W_in[:D], b_in[:D]) * (hd ** -0.5)
    k = linear(key, W_in[D:2 * D], b_in[D:2 * D])
    v = linear(value, W_in[2 * D:], b_in[2 * D:])
    q = q.contiguous().view(T, B * H, hd).transpose(0, 1)
    k = k.contiguous().view(-1, B * H, hd).transpose(0, 1)
    v = v.contiguous().view(-1, B * H, hd).transpose(0, 1)
    S = k.size(1)
    w = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        w = w.masked_fill(attn_mask.bool().unsqueeze(0), float('-inf'))
    if key_padding_mask is not None:
        w = w.view(B, H, T, S).masked_fill(key_padding_mask.bool().transpose(0, 1).unsqueeze(1).unsqueeze(2),
                                           float('-inf')).view(B * H, T, S)
    w = F.softmax(w, dim=-1)
    if weights_dropout:
        w = F.dropout(w, p=p, training=p > 0)
    attn = torch.bmm(w, v)
    if not weights_dropout:
        attn = F.dropout(attn, p=p, training=p > 0)
    attn = attn.transpose(0, 1).contiguous().view(T, B, D)
    out = linear(attn, W_out, b_out)
    if need_weights:
        return out, w.view(B, H, T, S).max(dim=1)[0].transpose(0, 1)
    return out, None


# SURVEY 8 f-0 forward half: gtos_rel_attn_banked_fwd / gtos_rel_grad_banked instead of the P-row tcgen05 kernels when the
# relation arrives factorised.  GTOS_BANKED_FWD=0 keeps the dense bf16 gather + gtos_rel_score / gtos_rel_grad.
_banked_fwd = os.environ.get("GTOS_BANKED_FWD", "1") == "1"
# dense relation tensor: projection GEMM + scores + softmax + dropout + P.V as ONE kernel (gtos_rel_attn_fwd) where a tile
# can hold all keys of its queries.  Built, parity-tested (tests/test_gpu_fused_dense.py), and measured: 9.25 ms vs 9.09 ms
# per config-2 step for gtos_rel_score + gtos_attn_fwd (same box) - the score kernel's epilogue warps are its bottleneck
# and the softmax / P.V tail lands on them.  So it is opt-in: GTOS_REL_FUSED_FWD=1 (read once, here and by the tile chooser).
_rel_fused_fwd = os.environ.get("GTOS_REL_FUSED_FWD", "0") == "1"


def banked_fwd_supported(banked, N, B, D, H):
    if not _banked_fwd or banked is None or banked.multi:
        return False
    if D not in (128, 256, 512) or 32 % H != 0 or D % H != 0:
        return False
    npad = (N + 3) // 4 * 4
    return 4 * (4 * H * npad + 8 * 4 * D) <= 200 * 1024 and tuple(banked.idx.shape) == (N, N, B)


class RelGradAcc:
    """holds the shared d_relation buffer of one GraphTransformer pass (dense relation), or - when the relation is a
    BankedRelation - the shared d_bank buffer and the segment-sum scratch"""

    def __init__(self, banked=None, n_layers=1):
        self.buf = None
        self.banked = banked
        self.n_layers = n_layers
        self.S = self.Wcat = self.spill = None
        self.slot = 0


class BankedRelation:
    """relation = bank[idx] kept factorised (caller generator/generator.py:76-79).

    bank [R,D] fp32 (output of RelationEncoder, may require grad), idx [N,N,B] int64 with idx[j][i][b] = bank row of
    the path i -> j (data.py:164-176).  The forward still feeds the dense bf16 tensor to the fused tcgen05 score
    kernel; the backward uses the bank structure: d bank and d relation_in_proj come from R-row GEMMs after one
    segmented sum of the per-pair gradients, so no fp32 [N,N,B,D] tensor (forward or backward) ever exists.
    Pass it as the `relation` argument of gtos_b200.GraphTransformer."""

    def __init__(self, bank, idx):
        _need_cuda(bank, idx)
        if bank.dim() != 2 or idx.dim() not in (3, 4) or idx.shape[0] != idx.shape[1]:
            raise ValueError(f"BankedRelation: bank [R,D] and idx [N,N,B] (or [N,N,B,K], evaluation batches) expected, "
                             f"got {tuple(bank.shape)}, {tuple(idx.shape)}")
        self.bank = bank
        self.idx = idx.contiguous()
        # evaluation batches carry up to K shortest paths per pair, 0 = empty slot (data.py:176-225); the relation of a
        # pair is the mean of their encodings (generator.py:83-88) - gather + mean fused, bf16 operand only
        self.multi = idx.dim() == 4
        with torch.no_grad():
            self._bank_c = bank.detach().contiguous()
            self.bankb = cast_bf16(self._bank_c)
        self._relb = None
        self.keys = self.order = None
        self._heads = None

    @property
    def relb(self):
        """the dense bf16 operand [N,N,B,D] of the tcgen05 score kernel - gathered on first use (the fused banked forward
        never needs it)"""
        if self._relb is None:
            idx, D = self.idx, self.bank.shape[1]
            with torch.no_grad():
                self._relb = torch.empty(*idx.shape[:3], D, dtype=torch.bfloat16, device=self.bank.device)
                if self.multi:
                    K = idx.shape[3]
                    _lib.check(_lib.load().gtos_bank_gather_mean(_p(self._bank_c), _p(idx), idx.numel() // K, K, D, None,
                                                                 _p(self._relb), _st()), "bank_gather_mean")
                else:
                    _lib.check(_lib.load().gtos_bank_gather(_p(self._bank_c), _p(idx), idx.numel(), D, None, _p(self._relb),
                                                            _st()), "bank_gather")
        return self._relb

    @property
    def requires_grad(self):
        return self.bank.requires_grad

    @property
    def shape(self):
        return (*self.idx.shape[:3], self.bank.shape[1])

    def prepare(self, H):
        """sort the G rows (tile-major pair rows of gtos_rel_grad) by bank row - once per batch"""
        if self._heads == H:
            return
        N, _, B = self.idx.shape
        R, D = self.bank.shape
        tiles = rel_tiling(N, B, D, H)["tiles"]
        raw = torch.empty(tiles * 128, dtype=torch.int32, device=self.idx.device)
        _lib.check(_lib.load().gtos_rel_pair_keys(_p(self.idx), N, B, D, H, R, _p(raw), _st()), "rel_pair_keys")
        keys, order = torch.sort(raw)
        self.keys, self.order = keys.contiguous(), order.to(torch.int32).contiguous()
        self._heads = H

    def dense(self):
        """the fp32 tensor the reference would build - bank[idx] (generator.py:79), or for evaluation batches the mean
        over a pair's paths with row 0 zeroed (generator.py:83-88) - for callers that need it materialised"""
        if not self.multi:
            return self.bank[self.idx]
        bank0 = torch.cat([torch.zeros_like(self.bank[:1]), self.bank[1:]], 0)
        cnt = self.idx.ne(0).sum(dim=3).clamp_(min=1).unsqueeze(-1).to(self.bank.dtype)
        return bank0[self.idx].sum(dim=3) / cnt


class BankTokenFn(torch.autograd.Function):
    """bank -> scalar token consumed by every layer's RelAttnFn; its backward runs after all of them and returns
    the d_bank they accumulated"""

    @staticmethod
    def forward(ctx, bank, acc):
        ctx.acc = acc
        ctx.set_materialize_grads(False)
        return bank.new_zeros(())

    @staticmethod
    def backward(ctx, _g):
        acc = ctx.acc
        if acc.S is None:
            return None, None
        d, _ = gemm_tn(acc.S, acc.Wcat, acc.Wcat.shape[0])          # [R, L*2D] x [D, L*2D]^T -> d bank [R, D]
        acc.S = acc.Wcat = acc.spill = None
        return d, None


class RelTokenFn(torch.autograd.Function):
    """relation -> scalar token consumed by every layer's RelAttnFn; its backward runs after all of them and returns
    the gradient they accumulated in `acc.buf` (one [N,N,B,D] tensor instead of L tensors + L-1 adds)."""

    @staticmethod
    def forward(ctx, relation, acc):
        ctx.acc = acc
        ctx.set_materialize_grads(False)
        return relation.new_zeros(())

    @staticmethod
    def backward(ctx, _g):
        buf, ctx.acc.buf = ctx.acc.buf, None
        return buf, None


# --------------------------------------------------------------------------------------------
# vanilla multi-head attention (transformer.py:98-173)
# --------------------------------------------------------------------------------------------
def attn_bwd_decoder(lib, d):
    """gtos_attn_bwd in decoder mode (dq, dk, dv all wanted).  Where one CTA holds every query row of a (batch, head) - all
    of gtos's decoder shapes except the 1-head alignment attention - the query-side kernel also produces dK, and the
    key side is left with dV = Pd^T dO, which needs nothing from the query side: it runs on a second stream BESIDE it
    (the dependent chain of a decoder attention backward is one kernel instead of two).  Returns the fork to join before
    dv is read."""
    if _side_enabled and lib.gtos_attn_bwd_dk_on_query_side(C.byref(d)) == 1:
        dq, dk, dqb, dkb = d.dq, d.dk, d.dq_bf16, d.dk_bf16
        with fork(which=4) as f_dv:
            d.bwd_part, d.dq, d.dk, d.dq_bf16, d.dk_bf16 = 2, None, None, None, None
            _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(dec, dV)")
        d.bwd_part, d.dq, d.dk, d.dq_bf16, d.dk_bf16 = 1, dq, dk, dqb, dkb
        _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(dec, dS dq dK)")
        return f_dv
    d.bwd_part = 0
    _lib.check(lib.gtos_attn_bwd(C.byref(d), _st()), "attn_bwd(dec)")
    return _NoFork()


class MHAFn(torch.autograd.Function):
    """query [T,B,D]; key [S,B,D] (value is key); returns (out, weights[B,H,T,S] | None)."""

    @staticmethod
    def forward(ctx, query, qb, key, kb, self_attn, key_pad, attn_mask, W_in, b_in, W_out, b_out, H, p,
                weights_dropout, need_weights):
        _need_cuda(query, key, W_in)
        lib = _lib.load()
        T, B, D = query.shape
        S = key.shape[0]
        hd = D // H
        dev = query.device
        qb2 = (qb if qb is not None else cast_bf16(query.contiguous().view(T * B, D))).view(T * B, _up8(D))
        Wib, Wit = weight_prep(W_in)
        Wob, Wot = weight_prep(W_out)
        if self_attn:
            kb2 = qb2
            proj, _ = gemm_tn(qb2, Wib, 3 * D, bias=b_in)                          # [TB, 3D]
            qp, kp, vp, ldq, ldk = proj.data_ptr(), proj.data_ptr() + 4 * D, proj.data_ptr() + 8 * D, 3 * D, 3 * D
            keep = (proj,)
        else:
            key2 = key.contiguous().view(S * B, D)
            kb2 = kb.view(S * B, _up8(D)) if kb is not None else torch.empty(S * B, _up8(D), dtype=torch.bfloat16, device=dev)
            pkv = torch.empty(S * B, 2 * D, dtype=torch.float32, device=dev)
            with fork() as f_kv:                    # the memory side (cast + K/V projection) runs beside the query projection
                if kb is None:
                    _lib.check(lib.gtos_cast_bf16(_p(key2), D, _p(kb2), _up8(D), S * B, D, _st()), "cast_bf16")
                _lib.check(lib.gtos_gemm_tn(_p(kb2), kb2.stride(0), Wib.data_ptr() + 2 * D * Wib.stride(0), Wib.stride(0),
                                            b_in.data_ptr() + 4 * D, _p(pkv), 2 * D, None, 0, S * B, 2 * D, _up8(D), 0, 0,
                                            _st()), "gemm_tn(kv)")
            pq, _ = gemm_tn(qb2, Wib, D, bias=b_in[:D], M=T * B)                    # rows 0..D of W_in
            f_kv.join()
            qp, kp, vp, ldq, ldk = pq.data_ptr(), pkv.data_ptr(), pkv.data_ptr() + 4 * D, D, 2 * D
            keep = (pq, pkv)
        probs = torch.empty(B, H, T, S, dtype=torch.float32, device=dev)
        p_w = p if weights_dropout else 0.0
        wts = torch.empty(B, H, T, S, dtype=torch.float32, device=dev) if need_weights else None
        att = torch.empty(T * B, D, dtype=torch.float32, device=dev)
        attb = torch.empty(T * B, _up8(D), dtype=torch.bfloat16, device=dev)
        seed, off = (rng_state(dev), new_seed_off()) if p > 0 else (None, 0)
        d = _attn_desc(T, S, B, H, hd)
        d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = qp, ldq, kp, ldk, vp, ldk
        d.scale, d.p_drop = hd ** -0.5, p_w
        d.key_pad, d.attn_mask = _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs, d.probs_dropped = _p(probs), _p(wts)
        d.out, d.ldo = _p(att), D
        d.out_bf16 = _p(attb) if (weights_dropout or p == 0) else None
        _lib.check(lib.gtos_attn_fwd(C.byref(d), _st()), "attn_fwd(dec)")
        off2 = 0
        if not weights_dropout and p > 0:                                          # transformer.py:156-157
            off2 = new_seed_off()
            dropout_f32(att, p, seed, off2, out=att)
            attb = cast_bf16(att)
        out, _ = gemm_tn(attb, Wob, D, bias=b_out)
        ctx.save_for_backward(qb2, kb2, probs, attb, Wit, Wot, key_pad, attn_mask, *keep)
        ctx.meta = (T, S, B, D, H, p, p_w, seed, off, off2, self_attn)
        ctx.set_materialize_grads(False)            # unused attention weights: no zero-filled [B,H,T,S] gradient
        if wts is None:
            return out.view(T, B, D), None
        return out.view(T, B, D), wts

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dwts):
        qb2, kb2, probs, attb, Wit, Wot, key_pad, attn_mask, *keep = ctx.saved_tensors
        T, S, B, D, H, p, p_w, seed, off, off2, self_attn = ctx.meta
        lib = _lib.load()
        hd = D // H
        if dout is None:
            dout = torch.zeros(T, B, D, dtype=torch.float32, device=qb2.device)
        dev = dout.device
        dout2 = dout.contiguous().view(T * B, D)
        doutb, db_out, f_db = grad_operand(dout, dout2)
        dW_out = torch.empty(D, D, dtype=torch.float32, device=dev)
        with fork() as f_out:
            gemm_nn(doutb, attb, D, D, out=dW_out)
        datt, _ = gemm_tn(doutb, Wot, D)
        if off2:
            dropout_f32(datt, p, seed, off2, out=datt)
        d = _attn_desc(T, S, B, H, hd)
        copies = _grad_copies and D % 8 == 0
        if self_attn:
            (proj,) = keep
            dproj = torch.empty(T * B, 3 * D, dtype=torch.float32, device=dev)
            dproj_b = torch.empty(T * B, 3 * D, dtype=torch.bfloat16, device=dev) if copies else None
            if copies:
                d.dq_bf16, d.dk_bf16, d.dv_bf16 = dproj_b.data_ptr(), dproj_b.data_ptr() + 2 * D, dproj_b.data_ptr() + 4 * D
            d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = (proj.data_ptr(), 3 * D, proj.data_ptr() + 4 * D, 3 * D,
                                                  proj.data_ptr() + 8 * D, 3 * D)
            d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = (dproj.data_ptr(), 3 * D, dproj.data_ptr() + 4 * D, 3 * D,
                                                        dproj.data_ptr() + 8 * D, 3 * D)
        else:
            pq, pkv = keep
            dpq = torch.empty(T * B, D, dtype=torch.float32, device=dev)
            dpkv = torch.empty(S * B, 2 * D, dtype=torch.float32, device=dev)
            dpq_b = torch.empty(T * B, D, dtype=torch.bfloat16, device=dev) if copies else None
            dpkv_b = torch.empty(S * B, 2 * D, dtype=torch.bfloat16, device=dev) if copies else None
            if copies:
                d.dq_bf16, d.dk_bf16, d.dv_bf16 = dpq_b.data_ptr(), dpkv_b.data_ptr(), dpkv_b.data_ptr() + 2 * D
            d.q, d.ldq, d.k, d.ldk, d.v, d.ldv = pq.data_ptr(), D, pkv.data_ptr(), 2 * D, pkv.data_ptr() + 4 * D, 2 * D
            d.dq, d.lddq, d.dk, d.lddk, d.dv, d.lddv = (dpq.data_ptr(), D, dpkv.data_ptr(), 2 * D,
                                                        dpkv.data_ptr() + 4 * D, 2 * D)
        ds_ts = torch.empty(B, H, T, S, dtype=torch.float32, device=dev)
        d.scale, d.p_drop = hd ** -0.5, p_w
        d.key_pad, d.attn_mask = _p(key_pad), _p(attn_mask)
        d.seed_ptr, d.seed_off = _p(seed), off
        d.probs = _p(probs)
        d.dout, d.lddo = _p(datt), D
        d.dprobs_extra = _p(dwts.contiguous()) if dwts is not None else None
        d.dscores_ts = _p(ds_ts)
        f_dv = attn_bwd_decoder(lib, d)
        dW_in = torch.empty(3 * D, D, dtype=torch.float32, device=dev)
        f_dv.join()                                 # dV (key side) ran beside dS / dq / dK
        if self_attn:
            dprojb, db_in, f_dbin = operand_or_cast(dproj, dproj_b)
            with fork() as f_in:
                gemm_nn(dprojb, qb2, 3 * D, D, out=dW_in)
            dq_in, _ = gemm_tn(dprojb, Wit, D)
            dk_in = None
        else:
            dpqb, dbq, f_dbin = operand_or_cast(dpq, dpq_b)
            dpkvb, dbkv, f_dbkv = operand_or_cast(dpkv, dpkv_b)
            f_dbin.join()
            f_dbkv.join()
            db_in = torch.cat([dbq, dbkv])
            with fork() as f_in:
                gemm_nn(dpqb, qb2, D, D, out=dW_in[:D])
                gemm_nn(dpkvb, kb2, 2 * D, D, out=dW_in[D:])
            dq_in, _ = gemm_tn(dpqb, Wit, D, K=D)                                   # Wt[:, :D]
            dk_in, _ = gemm_tn(dpkvb, Wit, D, K=2 * D, b_off=D)                     # Wt[:, D:3D]
            dk_in = dk_in.view(S, B, D)
        f_out.join_param(doutb, attb)
        f_in.join_param(qb2, kb2, *(([dprojb] if self_attn else [dpqb, dpkvb])))
        f_db.join()
        f_dbin.join_param(dproj if self_attn else dpq)
        return (dq_in.view(T, B, D), None, dk_in, None, None, None, None, dW_in, db_in, dW_out, db_out, None, None,
                None, None)


# --------------------------------------------------------------------------------------------
# RelationEncoder: embedding -> packed 2-layer bidirectional GRU -> Linear  (encoder.py:90-119)
# --------------------------------------------------------------------------------------------
def _up64(n):
    return (n + 63) // 64 * 64


def gru_row_counts(lengths, Lmax):
    """[#(len > t) for t in range(Lmax)] as Python ints - one small device-to-host read (the reference reads ALL lengths
    back here: `src_lengths.tolist()`, encoder.py:99).  Callers that know the lengths on the host (the data loader
    builds them there) pass the counts in instead and keep the step free of host syncs."""
    t = torch.arange(Lmax, device=lengths.device).unsqueeze(1)
    return [int(v) for v in (lengths.unsqueeze(0) > t).sum(1).tolist()]


class GRUBankFn(torch.autograd.Function):
    """tokens [Lmax,R] int64 (0 = pad), lengths [R] int64 -> [R, embed_dim].  `weights` is the flat list
    (w_ih, w_hh, b_ih, b_hh) per (layer, direction) in nn.GRU order.

    Forward: per (layer, direction, time step) ONE tcgen05 GEMM [x_t | h] x Wcat^T whose epilogue does the gate
    math, packed-sequence masking and writes h (fp32 + bf16 operand copy), the layer output and the saved gates
    (gtos_gru_step_fwd) - no gi / gh round trips through HBM.  Backward: BPTT with a gate kernel + accumulate
    GEMM per step, then three big GEMMs per (layer, direction) for dW_ih, dW_hh and dx.

    `counts` (list of Lmax ints, counts[t] = number of paths longer than t, or None): like pack_padded_sequence
    (encoder.py:93-99) the paths are sorted by length (on the device, stable) so that the live rows of time step t are the
    first counts[t] rows, and every per-step kernel runs on that prefix only - at config 2 (lengths 1..4) 21 % fewer
    row-steps, at translator settings (lengths 1..8) 37 %.  None: every row at every step (masked in the kernels)."""

    @staticmethod
    def forward(ctx, tokens, lengths, embed_w, out_w, out_b, num_layers, hidden, p, counts, *weights):
        _need_cuda(tokens, lengths, embed_w)
        lib = _lib.load()
        dev = embed_w.device
        Lmax, R = tokens.shape
        E = embed_w.shape[1]
        Hh = hidden
        rows = Lmax * R
        tokens = tokens.contiguous()
        lengths = lengths.contiguous()
        order = inv = None
        if counts is not None:
            counts = [min(int(c), R) for c in counts] + [0]
            if len(counts) != Lmax + 1 or any(counts[t] < counts[t + 1] for t in range(Lmax)) or (R > 0 and counts[0] > R):
                raise ValueError(f"GRUBankFn: row counts {counts[:-1]} are not a non-increasing list of {Lmax} values <= {R}")
            if all(c == R for c in counts[:-1]):
                counts = None                                                      # nothing to skip
        if counts is not None:
            # longest first; 8-bit keys = ONE radix pass (path lengths are <= 8 labels, data.py:153; int64 keys take 8)
            key = lengths.clamp(max=255).to(torch.uint8) if Lmax <= 255 else lengths
            order = torch.sort(key, descending=True, stable=True).indices
            inv = torch.empty_like(order)
            inv[order] = torch.arange(R, device=dev)
            tokens = tokens.index_select(1, order).contiguous()
            lengths = lengths.index_select(0, order).contiguous()
        n_at = (lambda t: counts[t]) if counts is not None else (lambda t: R)
        seed = rng_state(dev) if p > 0 else None
        off_e = new_seed_off() if p > 0 else 0
        xb = torch.empty(rows, _up8(E), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.gtos_embed_gather(_p(embed_w), _p(tokens), rows, E, None, _p(xb), _up8(E), p, _p(seed), off_e,
                                         _st()), "embed_gather")
        saved, layer_offs = [], []
        finals_b = torch.empty(R, 2 * Hh, dtype=torch.bfloat16, device=dev)
        Kin = E
        for l in range(num_layers):
            outb = torch.empty(rows, 2 * Hh, dtype=torch.bfloat16, device=dev) if l < num_layers - 1 else None
            clear = []                                                             # ranges to zero: ONE launch per layer
            if outb is not None and counts is not None:
                for t in range(Lmax):                                              # rows no kernel writes: zero (they are
                    if counts[t] < R:                                              # operands of the next layer's GEMMs)
                        clear.append(outb[t * R + counts[t]:(t + 1) * R])
            Kx = _up64(Kin)
            ldw = Kx + _up8(Hh)
            per_dir = []
            for d in range(2):                                                     # operands of both directions first
                w_ih, w_hh, b_ih, b_hh = weights[(l * 2 + d) * 4:(l * 2 + d) * 4 + 4]
                _, Wih_t = weight_prep(w_ih, want_b=False, plan=False)                         # for dx in backward
                _, Whh_t = weight_prep(w_hh, want_b=False, plan=False)                         # for dh in backward
                Wcat = torch.empty(4 * Hh, ldw, dtype=torch.bfloat16, device=dev)
                bcat = torch.empty(4 * Hh, dtype=torch.float32, device=dev)
                _lib.check(lib.gtos_gru_weight_prep(_p(w_ih.detach()), _p(w_hh.detach()), _p(b_ih.detach()),
                                                    _p(b_hh.detach()), Kin, Hh, Kx, _p(Wcat), ldw, _p(bcat), _st()),
                           "gru_weight_prep")
                # state indexed by processing step s (time t = s forward, Lmax-1-s reverse): hs[s] -> hs[s+1]
                gates = torch.empty(Lmax, R, 4 * Hh, dtype=torch.bfloat16, device=dev)
                hs = torch.empty(Lmax + 1, R, Hh, dtype=torch.float32, device=dev)
                hsb = torch.empty(Lmax + 1, R, Hh, dtype=torch.bfloat16, device=dev)
                clear += [hs[0], hsb[0]]
                if counts is not None:
                    # rows of hs / hsb[s] that step s - 1 does not write: a path that becomes live at step s (reverse
                    # direction) starts from h = 0, and hsb[s] as a whole is an operand of the dW_hh GEMM
                    for s_ in range(1, Lmax):
                        wrote = counts[s_ - 1] if d == 0 else counts[Lmax - s_]
                        if wrote < R:
                            clear.append(hsb[s_, wrote:])
                            if d == 1:
                                clear.append(hs[s_, wrote:counts[Lmax - 1 - s_]])
                per_dir.append((Wcat, bcat, gates, hs, hsb))
                saved += [xb, gates, hs, hsb, Wih_t, Whh_t]
            zero_regions(clear)

            def run_dir(d, xb=xb, outb=outb, Kin=Kin, Kx=Kx, ldw=ldw, per_dir=per_dir, last=(l == num_layers - 1)):
                Wcat, bcat, gates, hs, hsb = per_dir[d]
                for s in range(Lmax):
                    t = s if d == 0 else Lmax - 1 - s
                    n = n_at(t)
                    if n == 0:
                        continue
                    x_t = xb[t * R:(t + 1) * R]
                    out_t = outb[t * R:(t + 1) * R, d * Hh:(d + 1) * Hh] if not last else None
                    _lib.check(lib.gtos_gru_step_fwd(_p(x_t), x_t.stride(0), Kin, _p(hsb[s]), Hh, _p(hs[s]), _p(Wcat), ldw,
                                                     Kx, _p(bcat), _p(lengths), t, _p(hs[s + 1]), _p(hsb[s + 1]), Hh,
                                                     _p(out_t), 2 * Hh, _p(gates[s]), 4 * Hh, n, Hh, _st()),
                               "gru_step_fwd")
                if last:
                    if counts is None or d == 1:
                        finals_b[:, d * Hh:(d + 1) * Hh].copy_(hsb[Lmax])
                    else:
                        for t in range(Lmax):                                      # a path's final state: its last live step
                            lo, hi = counts[t + 1], counts[t]
                            if hi > lo:
                                finals_b[lo:hi, :Hh].copy_(hsb[t + 1, lo:hi])

            with fork(_gru_streams) as f_rev:                                      # reverse direction on the second stream
                run_dir(1)
            run_dir(0)
            f_rev.join()
            off_l = 0
            if l < num_layers - 1 and p > 0:                                       # nn.GRU inter-layer dropout
                off_l = new_seed_off()
                _lib.check(lib.gtos_dropout_bf16(_p(outb), outb.numel(), p, _p(seed), off_l, _st()), "dropout_bf16")
            layer_offs.append(off_l)
            xb = outb
            Kin = 2 * Hh
        if inv is not None:
            finals_b = finals_b.index_select(0, inv)                               # back to the caller's row order
        Wo_b, Wo_t = weight_prep(out_w)
        out, _ = gemm_tn(finals_b, Wo_b, out_w.shape[0], bias=out_b)
        ctx.save_for_backward(tokens, lengths, finals_b, Wo_t, *saved)
        ctx.meta = (Lmax, R, E, Hh, num_layers, p, seed, off_e, layer_offs, out_w.shape[0], embed_w.shape[0], counts, order)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        tokens, lengths, finals_b, Wo_t, *saved = ctx.saved_tensors
        Lmax, R, E, Hh, num_layers, p, seed, off_e, layer_offs, Dout, V, counts, order = ctx.meta
        lib = _lib.load()
        dev = dout.device
        rows = Lmax * R
        dout = dout.contiguous()
        doutb, db_out = cast_colsum(dout)
        dW_out = gemm_nn(doutb, finals_b, Dout, 2 * Hh)
        dfinals, _ = gemm_tn(doutb, Wo_t, 2 * Hh)                                   # [R, 2H]
        if order is not None:
            dfinals = dfinals.index_select(0, order)                               # into the length-sorted row order
        n_at = (lambda t: counts[t]) if counts is not None else (lambda t: R)
        wgrads = [None] * (num_layers * 8)
        d_layer_out = None                                                         # [rows, 2H] fp32
        for l in range(num_layers - 1, -1, -1):
            dgi_cat = torch.empty(rows, 6 * Hh, dtype=torch.bfloat16, device=dev)      # [dgi_fwd | dgi_rev], time order
            Wih_t_cat = []
            per_dir = []
            clear = []                                                                 # ranges to zero: ONE launch per layer
            dh0 = []                                                                   # initial dh of each direction
            for d in range(2):                                                         # outputs of both directions first
                xb, gates, hs, hsb, Wih_t, Whh_t = saved[(l * 2 + d) * 6:(l * 2 + d) * 6 + 6]
                Kin = Wih_t.shape[0]
                base = (l * 2 + d) * 4
                wgrads[base + 0] = torch.empty(3 * Hh, Kin, dtype=torch.float32, device=dev)
                wgrads[base + 1] = torch.empty(3 * Hh, Hh, dtype=torch.float32, device=dev)
                wgrads[base + 2] = torch.empty(3 * Hh, dtype=torch.float32, device=dev)
                wgrads[base + 3] = torch.empty(3 * Hh, dtype=torch.float32, device=dev)
                clear += [wgrads[base + 2], wgrads[base + 3]]
                dgh = torch.empty(rows, 3 * Hh, dtype=torch.bfloat16, device=dev)      # rows in step order s
                if counts is not None:
                    for s_ in range(Lmax):                                             # rows the gate kernel skips: zero,
                        n_ = counts[s_ if d == 0 else Lmax - 1 - s_]                   # they are operands of the dW GEMMs
                        if n_ < R:
                            clear.append(dgh[s_ * R + n_:(s_ + 1) * R])
                per_dir.append(dgh)
                Wih_t_cat.append(Wih_t)
                if l == num_layers - 1:
                    dh0.append(None)
                else:
                    dh0.append(torch.empty(R, Hh, dtype=torch.float32, device=dev))
                    clear.append(dh0[-1])
            if counts is not None:
                for t in range(Lmax):
                    if counts[t] < R:
                        clear.append(dgi_cat[t * R + counts[t]:(t + 1) * R])
            zero_regions(clear)

            def run_dir(d, l=l, per_dir=per_dir, dgi_cat=dgi_cat, d_layer_out=d_layer_out, dh0=dh0):
                xb, gates, hs, hsb, Wih_t, Whh_t = saved[(l * 2 + d) * 6:(l * 2 + d) * 6 + 6]
                Kin = Wih_t.shape[0]
                base = (l * 2 + d) * 4
                dgh, db_ih, db_hh = per_dir[d], wgrads[base + 2], wgrads[base + 3]
                # ONE dh buffer updated in place on the live prefix: a path whose last live step is still to come (in
                # this backward order) keeps the gradient of its final state until then
                if l == num_layers - 1:
                    dh = dfinals[:, d * Hh:(d + 1) * Hh].contiguous()
                else:
                    dh = dh0[d]                                                        # zeroed with the layer's other ranges
                dh_part = torch.empty_like(dh)                                         # dh * z (pass-through for finished rows)
                dgi = dgi_cat[:, d * 3 * Hh:(d + 1) * 3 * Hh]                          # rows in time order t
                for s in range(Lmax - 1, -1, -1):
                    t = s if d == 0 else Lmax - 1 - s
                    n = n_at(t)
                    if n == 0:
                        continue
                    dout_t = d_layer_out[t * R:(t + 1) * R, d * Hh:(d + 1) * Hh] if d_layer_out is not None else None
                    dgi_t, dgh_s = dgi[t * R:(t + 1) * R], dgh[s * R:(s + 1) * R]
                    _lib.check(lib.gtos_gru_gate_bwd(_p(dh), _p(dout_t), dout_t.stride(0) if dout_t is not None else 0,
                                                     _p(gates[s]), _p(hs[s]), _p(lengths), t, _p(dh_part),
                                                     _p(dgi_t), dgi_t.stride(0), _p(dgh_s), 3 * Hh, _p(db_ih), _p(db_hh),
                                                     n, Hh, _st()), "gru_gate_bwd")
                    # dh <- dgh @ W_hh + dh * z   (rows of the prefix; the kernel reads dh_part and dgh only)
                    _lib.check(lib.gtos_gemm_tn_add(_p(dgh_s), 3 * Hh, _p(Whh_t), Whh_t.stride(0), None, _p(dh_part), Hh,
                                                    _p(dh), Hh, n, Hh, 3 * Hh, _st()), "gemm_tn_add")
                gemm_nn(dgi, xb, 3 * Hh, Kin, out=wgrads[base + 0])
                gemm_nn(dgh, hsb[:Lmax].view(rows, Hh), 3 * Hh, Hh, out=wgrads[base + 1])

            with fork(_gru_streams) as f_rev:                                          # reverse direction on the second stream
                run_dir(1)
            run_dir(0)
            f_rev.join()
            Kin = Wih_t_cat[0].shape[0]
            # dx of both directions in ONE GEMM: [dgi_fwd | dgi_rev] x [Wih_fwd^T ; Wih_rev^T]
            Wcat_t = torch.cat(Wih_t_cat, dim=1) if Wih_t_cat[0].shape[1] == 3 * Hh else None
            if Wcat_t is not None:
                dx, _ = gemm_tn(dgi_cat, Wcat_t, Kin)
            else:                                                                      # padded 3H: keep two GEMMs
                dx, _ = gemm_tn(dgi_cat[:, :3 * Hh], Wih_t_cat[0], Kin, K=3 * Hh)
                gemm_tn(dgi_cat[:, 3 * Hh:], Wih_t_cat[1], Kin, K=3 * Hh, out=dx, accumulate=True)
            if l > 0:
                if layer_offs[l - 1]:
                    dropout_f32(dx, p, seed, layer_offs[l - 1], out=dx)
                d_layer_out = dx
            else:
                d_embed = torch.zeros(V, E, dtype=torch.float32, device=dev)
                _lib.check(lib.gtos_embed_scatter_add(_p(dx), _p(tokens), rows, E, _p(d_embed), p, _p(seed), off_e,
                                                      _st()), "embed_scatter_add")
        return (None, None, d_embed, dW_out, db_out, None, None, None, None, *wgrads)


# --------------------------------------------------------------------------------------------
# plain Linear on the tcgen05 GEMM (TokenGenerator's transfer / vocabulary projection, decoder.py:39,44)
# --------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        _need_cuda(x, W)
        shape = x.shape
        K = shape[-1]
        N = W.shape[0]
        xb = cast_bf16(x.contiguous().view(-1, K))
        Wb, Wt = weight_prep(W)
        y, _ = gemm_tn(xb, Wb, N, bias=b)
        ctx.save_for_backward(xb, Wt)
        ctx.meta = (shape, N, K, b is not None)
        return y.view(*shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xb, Wt = ctx.saved_tensors
        shape, N, K, has_b = ctx.meta
        dy2 = dy.contiguous().view(-1, N)
        dyb, db, f_db = grad_operand(dy, dy2)
        dW = torch.empty(N, K, dtype=torch.float32, device=dy.device)
        with fork() as f:
            gemm_nn(dyb, xb, N, K, out=dW)
        db = db if has_b else None
        dx, _ = gemm_tn(dyb, Wt, K)
        f.join_param(dyb, xb)
        f_db.join()
        return dx.view(shape), dW, db


def linear(x, W, b=None):
    if _precision == "fp32":
        from . import ops32
        return ops32.Linear32Fn.apply(x, W, b)
    return LinearFn.apply(x, W, b)


# --------------------------------------------------------------------------------------------
# bank -> dense relation gather (generator.py:79) with the bf16 operand copy made in the same pass
# --------------------------------------------------------------------------------------------
_bank_sorted_bwd = os.environ.get("GTOS_BANK_SORTED_BWD", "1") == "1"


class BankGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bank, idx):
        _need_cuda(bank, idx)
        bank = bank.contiguous()
        idx = idx.contiguous()
        R, D = bank.shape
        P = idx.numel()
        rel = torch.empty(*idx.shape, D, dtype=torch.float32, device=bank.device)
        relb = torch.empty(*idx.shape, D, dtype=torch.bfloat16, device=bank.device)
        _lib.check(_lib.load().gtos_bank_gather(_p(bank), _p(idx), P, D, _p(rel), _p(relb), _st()), "bank_gather")
        ctx.sorted = None
        if _bank_sorted_bwd and ctx.needs_input_grad[0] and P > 0:
            # the backward sums d_rel rows per bank row: sort the pairs by bank row now, beside the encoder's forward
            with fork(True, which=3) as f_sort:
                k32, order = torch.sort(idx.view(-1).to(torch.int32), stable=False)     # 4 radix passes instead of 8
                keys = k32.to(torch.int64)
            # joined right here: the sort (~40 us) runs beside the gather kernel (~70 us) and is over before it, and a fork
            # left open until the backward would break a CUDA-graph capture whose backward never reaches this node
            f_sort.join()
            ctx.sorted = True
            ctx.save_for_backward(idx, keys, order)
        else:
            ctx.save_for_backward(idx)
        ctx.meta = (R, D, P)
        ctx.mark_non_differentiable(relb)
        ctx.set_materialize_grads(False)            # no zero-filled [N,N,B,D] "gradient" for the bf16 copy
        return rel, relb

    @staticmethod
    @once_differentiable
    def backward(ctx, d_rel, _d_relb):
        idx = ctx.saved_tensors[0]
        R, D, P = ctx.meta
        if d_rel is None:
            return torch.zeros(R, D, dtype=torch.float32, device=idx.device), None
        d_rel = d_rel.contiguous()
        d_bank = torch.empty(R, D, dtype=torch.float32, device=d_rel.device)
        if ctx.sorted is not None:
            _, keys, order = ctx.saved_tensors
            _lib.check(_lib.load().gtos_bank_segsum(_p(d_rel), _p(order), _p(keys), P, D, _p(d_bank), R, _st()), "bank_segsum")
        else:
            _lib.check(_lib.load().gtos_bank_scatter_add(_p(d_rel), _p(idx), P, D, _p(d_bank), R, _st()), "bank_scatter_add")
        return d_bank, None


def bank_gather(bank, idx):
    """relation = bank[idx] as (fp32 dense tensor, bf16 copy).  The dense tensor carries the bf16 copy (it survives
    .view / .reshape / .contiguous, see GatheredRelation) so GraphTransformer.forward does not stage it again."""
    rel, relb = BankGatherFn.apply(bank, idx)
    rel = rel.as_subclass(GatheredRelation)
    rel._gtos_bf16 = (relb, rel._version)
    rel._gtos_src = (bank, idx, rel._version)       # provenance: this tensor IS bank[idx] while its version stands
    return rel


_VIEW_FUNCS = None


def _view_funcs():
    global _VIEW_FUNCS
    if _VIEW_FUNCS is None:
        T = torch.Tensor
        _VIEW_FUNCS = {T.view, T.reshape, T.contiguous, T.view_as, T.reshape_as, torch.reshape}
    return _VIEW_FUNCS


class GatheredRelation(torch.Tensor):
    """The dense fp32 relation tensor bank[idx] as ops.bank_gather builds it.  An ordinary tensor for every consumer; the
    only extra is that the bf16 operand copy made in the same pass rides along through shape-only views
    (`.view(*idx.size(), -1)`, generator.py:79), so the graph encoder does not have to re-read 4 bytes per element to
    re-make it.  The tag records the tensor's version counter: any in-place write invalidates it."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        if func in _view_funcs() and isinstance(out, torch.Tensor) and args and isinstance(args[0], GatheredRelation):
            src = args[0]
            tag = getattr(src, "_gtos_bf16", None)
            if (tag is not None and tag[1] == src._version and out.numel() == src.numel() and out.dtype == src.dtype
                    and out.is_contiguous() and out.data_ptr() == src.data_ptr()):
                if out is not src:
                    out = out.as_subclass(GatheredRelation)
                out._gtos_bf16 = (tag[0].view(out.shape), tag[1])
                prov = getattr(src, "_gtos_src", None)
                if prov is not None and prov[2] == src._version:
                    out._gtos_src = prov
        return out


def staged_relation_bf16(relation):
    """the bf16 operand copy a GatheredRelation carries, if it is still valid for `relation`"""
    tag = getattr(relation, "_gtos_bf16", None)
    if tag is None:
        return None
    relb, version = tag
    if version != relation._version or relb.shape != relation.shape or relb.device != relation.device:
        return None
    return relb


_bank_tensor = os.environ.get("GTOS_BANK_TENSOR", "1") == "1"
# GTOS_REL_PROVENANCE=1 (opt-in): a dense relation tensor that ops.bank_gather built - e.g. through the unchanged caller's
# index_select on a BankTensor - remembers (bank, idx).  While nobody has written to it, GraphTransformer may treat it as
# the factorised relation it is without any caller change: same function of (bank, idx), gradient delivered
# to the bank directly.  Off by default: the dense kernels are the contract the headline is measured on.
_rel_provenance = os.environ.get("GTOS_REL_PROVENANCE", "0") == "1"


def factorised_source(relation):
    """BankedRelation equivalent of a dense relation tensor with valid provenance, or None"""
    if not _rel_provenance:
        return None
    prov = getattr(relation, "_gtos_src", None)
    if prov is None or prov[2] != relation._version or relation.dim() != 4:
        return None
    bank, idx, _ = prov
    N1, N2, B, D = relation.shape
    if idx.numel() != N1 * N2 * B or N1 != N2 or bank.dim() != 2 or bank.shape[1] != D:
        return None
    return BankedRelation(bank, idx.view(N1, N2, B))


class BankTensor(torch.Tensor):
    """RelationEncoder's output [R, D].  The reference's caller turns it into the dense relation tensor with
    `relation.index_select(0, inp['relation'].view(-1)).view(*inp['relation'].size(), -1)` (generator.py:79;
    translator/generator.py:73).  That line keeps working unchanged: index_select along dim 0 of a BankTensor runs
    gtos_bank_gather (same fp32 values, plus the bf16 operand copy in the same pass) and its backward runs
    gtos_bank_scatter_add instead of ATen's index_add_ (CUPTI timeline at config 2: 478 us -> the <TL> row alone receives
    41 % of the pairs).  Every other operation sees a plain tensor.  GTOS_BANK_TENSOR=0 returns a plain tensor."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.Tensor.index_select, torch.index_select) and not kwargs.get("out"):
            bank = args[0] if args else kwargs.get("input")
            dim = args[1] if len(args) > 1 else kwargs.get("dim")
            index = args[2] if len(args) > 2 else kwargs.get("index")
            if True:
                if (isinstance(bank, BankTensor) and dim in (0, -2) and bank.dim() == 2 and bank.is_cuda
                        and torch.is_tensor(index) and index.dim() == 1 and index.dtype == torch.int64 and index.is_cuda
                        and bank.dtype == torch.float32 and bank.shape[1] % 4 == 0):
                    with torch._C.DisableTorchFunctionSubclass():
                        return bank_gather(bank.as_subclass(torch.Tensor), index)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)


def as_bank_tensor(bank):
    return bank.as_subclass(BankTensor) if (_bank_tensor and bank.is_cuda) else bank


# --------------------------------------------------------------------------------------------
# TokenGenerator training tail (decoder.py:42-64): fused copy/generate NLL over the vocabulary logits
# --------------------------------------------------------------------------------------------
class TokenNLLFn(torch.autograd.Function):
    """logits [T,B,V], gate_logits [T,B,2], align [T,B,S] fp32; copy_seq [S,B], target [T,B] int64 -> loss [T,B]."""

    @staticmethod
    def forward(ctx, logits, gate_logits, align, copy_seq, target, pad_idx):
        _need_cuda(logits, gate_logits, align)
        T, B, V = logits.shape
        S = align.shape[-1]
        logits, gate_logits, align = logits.contiguous(), gate_logits.contiguous(), align.contiguous()
        copy_seq, target = copy_seq.contiguous(), target.contiguous()
        rows = T * B
        loss = torch.empty(T, B, dtype=torch.float32, device=logits.device)
        stats = torch.empty(rows, 6, dtype=torch.float32, device=logits.device)
        _lib.check(_lib.load().gtos_token_nll_fwd(_p(logits), V, V, _p(gate_logits), _p(align), S, _p(copy_seq), _p(target),
                                                  rows, B, int(pad_idx), _p(loss), _p(stats), _st()), "token_nll_fwd")
        ctx.save_for_backward(logits, align, copy_seq, target, stats)
        ctx.meta = (T, B, V, S, int(pad_idx))
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        logits, align, copy_seq, target, stats = ctx.saved_tensors
        T, B, V, S, pad_idx = ctx.meta
        dev = dloss.device
        dloss = dloss.contiguous()
        dlogits = torch.empty(T, B, V, dtype=torch.float32, device=dev)
        dgate = torch.empty(T, B, 2, dtype=torch.float32, device=dev)
        dalign = torch.empty(T, B, S, dtype=torch.float32, device=dev)
        # the bf16 operand copy of d logits for the vocabulary projection's backward GEMMs is written in the same pass
        # (tagged on the gradient like AddLayerNormFn's, see grad_operand): no cast pass over the [T*B, V] tensor
        dlb = torch.empty(T * B, V, dtype=torch.bfloat16, device=dev) if (V % 8 == 0 and _grad_tags
                                                                          and _precision != "fp32") else None
        _lib.check(_lib.load().gtos_token_nll_bwd(_p(dloss), _p(logits), V, V, _p(align), S, _p(copy_seq), _p(target), T * B,
                                                  B, pad_idx, _p(stats), _p(dlogits), V, _p(dgate), _p(dalign), _p(dlb), V,
                                                  _st()), "token_nll_bwd")
        if dlb is not None:
            dlogits._gtos_bf16 = (dlb, dlogits._version)
        return dlogits, dgate, dalign, None, None, None


def token_nll(logits, gate_logits, align, copy_seq, target, pad_idx):
    return TokenNLLFn.apply(logits, gate_logits, align, copy_seq, target, pad_idx)


# --------------------------------------------------------------------------------------------
# TokenGenerator work=True tail (decoder.py:42-59): log-prob table over the batch-extended vocabulary, one kernel
# --------------------------------------------------------------------------------------------
def token_logprob(logits, gate_logits, align, copy_seq, src_index, width, B=None):
    """log-prob table [rows, width] (decoder.py:42-59).  logits [rows,V], gate_logits [rows,2], align [rows,S] fp32;
    copy_seq [S,Bsrc] int64; row -> graph through src_index (int32) or row % B."""
    _need_cuda(logits, gate_logits, align, copy_seq)
    rows, V = logits.shape
    S, Bsrc = copy_seq.shape
    logits, gate_logits, align = logits.contiguous(), gate_logits.contiguous(), align.contiguous()
    table = torch.empty(rows, width, dtype=torch.float32, device=logits.device)
    _lib.check(_lib.load().gtos_token_logprob(_p(logits), V, V, _p(gate_logits), _p(align), S, _p(copy_seq.contiguous()), Bsrc,
                                              _p(src_index), rows, B if B is not None else Bsrc, _p(table), width, width,
                                              _st()), "token_logprob")
    return table


def token_topk(logits, gate_logits, align, copy_seq, src_index, width, k, B=None, want_table=False):
    """the k best entries of every row of the log-prob table (generator.py:157), computed without materialising it:
    returns (values fp32 [rows,k], token ids int32 [rows,k]) (+ the table when want_table)."""
    _need_cuda(logits, gate_logits, align, copy_seq)
    rows, V = logits.shape
    S, Bsrc = copy_seq.shape
    logits, gate_logits, align = logits.contiguous(), gate_logits.contiguous(), align.contiguous()
    dev = logits.device
    top_val = torch.empty(rows, k, dtype=torch.float32, device=dev)
    top_idx = torch.empty(rows, k, dtype=torch.int32, device=dev)
    table = torch.empty(rows, width, dtype=torch.float32, device=dev) if want_table else None
    _lib.check(_lib.load().gtos_token_topk(_p(logits), V, V, _p(gate_logits), _p(align), S, _p(copy_seq.contiguous()), Bsrc,
                                           _p(src_index), rows, B if B is not None else Bsrc, width, k, _p(top_val),
                                           _p(top_idx), _p(table), width, _st()), "token_topk")
    return (top_val, top_idx, table) if want_table else (top_val, top_idx)
