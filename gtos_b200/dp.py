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
            torch.cat([p.grad.reshape(-1) for p in self.params], out=self.flat)

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
