"""The optimizer step of the reference's training loop on flat buffers.

Reference, per step (generator/train.py:148-154): `average_gradients` -> `clip_grad_norm_(model.parameters(), 1.0)` ->
`update_lr` (:81-83) -> `AdamWeightDecayOptimizer.step()` (generator/adam.py:28-87; two parameter groups, weight decay
1e-4 on everything that is not a bias or a LayerNorm parameter, train.py:123-132) -> `zero_grad()`: ~8 small kernels
for each of 182 parameters.  Here parameters, gradients and both moments live in four flat fp32 buffers laid out
[decayed parameters | undecayed parameters]; a step is `gtos_grad_sumsq` (two launches, deterministic) +
`gtos_adam_step` (one launch), and the learning rate is read from device memory so the step can sit inside a captured
CUDA graph.  The gradient buffer is the one `dp.FlatGradBucket` all-reduces.
"""
import torch

from . import _lib
from .dp import FlatGradBucket
from .ops import _need_cuda, _p, _st


def default_no_decay(name):
    """train.py:126: `name.endswith('bias') or 'layer_norm' in name`"""
    return name.endswith("bias") or "layer_norm" in name


def noam_lr(embed_size, steps, warmup_steps):
    """update_lr, train.py:81-83"""
    return embed_size ** -0.5 * min(steps ** -0.5, steps * (warmup_steps ** -1.5))


class FlatAdam:
    """AdamWeightDecayOptimizer(lr, betas=(0.9, 0.999), eps=1e-6) + clip_grad_norm_(max_norm) on flat buffers.

    named_params: iterable of (name, Parameter).  Every parameter's storage is MOVED into one flat buffer
    (`p.data` becomes a view), ordered decayed-first; `self.bucket` is the matching flat gradient bucket."""

    def __init__(self, named_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=1e-4, max_norm=1.0,
                 no_decay=default_no_decay, bind_grads=True):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("FlatAdam: no trainable parameters")
        decay = [p for n, p in named if not no_decay(n)]
        rest = [p for n, p in named if no_decay(n)]
        self.params = decay + rest
        _need_cuda(*self.params)
        dev = self.params[0].device
        self.n_decay = sum(p.numel() for p in decay)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.empty(self.numel, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                view = self.flat[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.bucket = FlatGradBucket(self.params, bind=bind_grads)
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, eps, weight_decay, max_norm
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = torch.empty(1, dtype=torch.float32).pin_memory()
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ws = torch.empty(int(_lib.load().gtos_grad_sumsq_workspace()), dtype=torch.float32, device=dev)
        self.steps = 0

    def set_lr(self, lr):
        """host -> device copy of the scheduled learning rate (outside a captured graph)"""
        self._lr_host[0] = float(lr)
        self.lr_dev.copy_(self._lr_host, non_blocking=True)

    def zero_grad(self):
        """use this instead of model.zero_grad(): it keeps every .grad a view of the flat buffer (a stray fresh .grad is
        still picked up by step(), at the cost of one copy per parameter).  Unlike the reference (adam.py:41-43, which
        skips parameters whose grad is None) a parameter that received no gradient still gets its weight decay and moment
        decay: the flat kernel has no per-parameter skip."""
        self.bucket.zero()

    def step(self):
        """clip + Adam over the flat buffers; enqueues 3 kernels on the current stream, no host sync.
        (Call bucket.all_reduce_mean() first when data parallel.)"""
        lib = _lib.load()
        self.bucket.pack()
        self.bucket.adopt()          # gradients written outside the flat views (zero_grad(set_to_none=True)) are copied in
        g = self.bucket.flat
        nsq = None
        if self.max_norm is not None:
            _lib.check(lib.gtos_grad_sumsq(_p(g), self.numel, _p(self.norm_sq), _p(self._ws), _st()), "grad_sumsq")
            nsq = self.norm_sq
        _lib.check(lib.gtos_adam_step(_p(self.flat), _p(g), _p(self.exp_avg), _p(self.exp_avg_sq), self.numel, self.n_decay,
                                      _p(self.lr_dev), self.betas[0], self.betas[1], self.eps, self.weight_decay, _p(nsq),
                                      float(self.max_norm or 0.0), _st()), "adam_step")
        self.steps += 1

    def grad_norm(self):
        """total gradient norm of the last step (device tensor; reading it syncs)"""
        return self.norm_sq.sqrt()
