"""Data-parallel gradient exchange: ONE all-reduce(mean) over a flat fp32 gradient bucket per step,
replacing the reference's 182 per-parameter all-reduce + divide calls (generator/train.py:74-79).

Every parameter's ``.grad`` is a view into one contiguous buffer, so backward writes gradients in
place and the collective needs no packing copy.  torch.distributed (NCCL over NVLink/NVSwitch on the
B200 box, gloo in the CPU tests) is the transport.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    """bind=True: every .grad is a view of the flat buffer (backward accumulates in place, no packing copy, but
    autograd then issues one small add per parameter).  bind=False: backward writes fresh gradients and
    pack() gathers them with ONE multi-tensor copy right before the collective - fewer launches per step."""

    def __init__(self, params, bind=True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.bound = bind
        self.numel = total
        if bind:
            off = 0
            for p in self.params:
                n = p.numel()
                p.grad = self.flat[off:off + n].view_as(p)
                off += n

    def zero(self):
        if self.bound:
            self.flat.zero_()
        else:
            for p in self.params:
                p.grad = None

    def pack(self):
        """gather the per-parameter gradients into the flat buffer (no-op when bound)"""
        if not self.bound:
            # a trainable parameter that took no part in this step has no gradient: it contributes zeros (the reference's
            # average_gradients skips it, train.py:76-77)
            torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params],
                      out=self.flat)

    def unpack(self):
        """scatter the (averaged) flat buffer back into the per-parameter gradients (no-op when bound)"""
        if not self.bound:
            off = 0
            for p in self.params:
                n = p.numel()
                p.grad = self.flat[off:off + n].view_as(p)
                off += n

    def rebind(self):
        """re-attach the views if something replaced .grad (e.g. zero_grad(set_to_none=True))."""
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * self.flat.element_size():
                p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def adopt(self):
        """bound buckets only: a backward pass that ran after `model.zero_grad()` (set_to_none=True is torch's default)
        wrote FRESH .grad tensors instead of accumulating into the views.  Copy those into the flat buffer and re-attach
        the views; a parameter without a gradient gets a zero slice.  Returns the number of parameters that had strayed."""
        if not self.bound:
            return 0
        off, strayed = 0, 0
        for p in self.params:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
                strayed += 1
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
                strayed += 1
            off += n
        return strayed

    def all_reduce_mean(self, group=None):
        """average_gradients (train.py:74-79) as one collective."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


def shard_range(total, rank, world):
    """contiguous equal shards of a global batch: graphs [lo, hi) for this rank."""
    if total % world != 0:
        raise ValueError(f"global batch {total} must divide evenly over {world} ranks (per-sequence loss mean, decoder.py:92-94)")
    per = total // world
    return rank * per, (rank + 1) * per


class OverlappedGradBuckets:
    """The flat gradient exchange split into a few buckets that are all-reduced WHILE the backward pass is still running
    ("optionally 2-3 buckets launched from backward hooks to hide it").

    groups: list of parameter lists in the order their gradients become final during backward (e.g. decoder + sentence
    encoder, graph encoder, relation encoder).  Each group owns a contiguous slice of one flat buffer.  A
    post-accumulate-grad hook counts the group's parameters down; when the last gradient of a group lands, the group is
    packed into its slice (one multi-tensor copy) and its all-reduce(mean) is issued asynchronously - the collective runs
    on the process group's own stream while the caller's stream continues with the rest of the backward.  `finish()`
    joins the outstanding collectives (and flushes groups some of whose parameters received no gradient).
    Everything is stream-ordered, so the whole step - hooks included - can be captured in one CUDA graph."""

    def __init__(self, groups, group=None):
        self.groups = [[p for p in g if p.requires_grad] for g in groups]
        self.groups = [g for g in self.groups if g]
        if not self.groups:
            raise ValueError("OverlappedGradBuckets: no trainable parameters")
        dev, dt = self.groups[0][0].device, self.groups[0][0].dtype
        self.numel = sum(p.numel() for g in self.groups for p in g)
        self.flat = torch.zeros(self.numel, dtype=dt, device=dev)
        self.slices, off = [], 0
        for g in self.groups:
            n = sum(p.numel() for p in g)
            self.slices.append(self.flat[off:off + n])
            off += n
        self.pg = group
        self._left = [0] * len(self.groups)
        self._done = [False] * len(self.groups)
        self._works = []
        self._hooks = []
        self.enabled = True
        for gi, g in enumerate(self.groups):
            for p in g:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(gi)))
        self.zero()

    def _make_hook(self, gi):
        def hook(_p):
            if not self.enabled:
                return
            self._left[gi] -= 1
            if self._left[gi] == 0:
                self._launch(gi)
        return hook

    def zero(self):
        """call before every backward: forget the gradients and re-arm the countdowns"""
        for gi, g in enumerate(self.groups):
            for p in g:
                p.grad = None
            self._left[gi] = len(g)
            self._done[gi] = False
        self._works = []

    def _launch(self, gi):
        g = self.groups[gi]
        if self.slices[gi].is_cuda:
            from . import ops
            ops.join_deferred_now()          # weight-gradient kernels still running on the side stream (ops.deferred_param_grads)
        parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in g]
        torch.cat(parts, out=self.slices[gi])
        self._done[gi] = True
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.pg) == 1:
            return
        if dist.get_backend(self.pg) == "nccl":
            self._works.append((gi, dist.all_reduce(self.slices[gi], op=dist.ReduceOp.AVG, group=self.pg, async_op=True), False))
        else:
            self._works.append((gi, dist.all_reduce(self.slices[gi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True), True))

    def finish(self):
        """after backward: flush groups that never completed, then make the current stream wait for every collective"""
        for gi in range(len(self.groups)):
            if not self._done[gi]:
                self._launch(gi)
        for gi, w, divide in self._works:
            w.wait()
            if divide:
                self.slices[gi].div_(dist.get_world_size(self.pg))
        self._works = []

    def unpack(self):
        """point every .grad at its (averaged) slice of the flat buffer"""
        for g, sl in zip(self.groups, self.slices):
            off = 0
            for p in g:
                n = p.numel()
                p.grad = sl[off:off + n].view_as(p)
                off += n

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
