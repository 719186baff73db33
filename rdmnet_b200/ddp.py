"""Data-parallel training plumbing (BASELINE config 4): the reference wraps its model in
torch DistributedDataParallel (geotransformer/engine/base_trainer.py:181-191: fp32 gradients, 25 MB buckets) and
all-reduces every logged scalar separately each iteration (utils/torch.py:16-21, base_trainer.py:236). Here:

  BucketedGradAllReduce   gradients are packed per bucket into a flat bf16 (or fp32) buffer as soon as the bucket's last
                          gradient has been accumulated (post-accumulate-grad hooks), the NCCL all-reduce of the bucket is
                          launched asynchronously - it overlaps the rest of the backward pass -, and `finish()` (before the
                          optimizer step) waits, averages and unpacks. 25.3 M parameters = 50.6 MB in bf16 per step.
  all_reduce_scalars      ONE collective for all logged scalars instead of one per key.

One process per GPU (torchrun); NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests. The hooks only pack,
cast and launch collectives: no arithmetic of the model runs here."""
import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, module, bucket_bytes=25 << 20, comm_dtype=torch.bfloat16, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.comm_dtype = comm_dtype
        params = [p for p in module.parameters() if p.requires_grad]
        # buckets in REVERSE registration order: gradients become ready roughly back to front
        self.buckets, cur, size = [], [], 0
        esz = torch.tensor([], dtype=comm_dtype).element_size()
        for p in reversed(params):
            cur.append(p)
            size += p.numel() * esz
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=comm_dtype, device=b[0].device) for b in self.buckets]
        self.where = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b:
                self.where[p] = (bi, off)
                off += p.numel()
        self.pending = [0] * len(self.buckets)
        self.handles = [None] * len(self.buckets)
        self.bytes_per_step = sum(f.numel() * f.element_size() for f in self.flat)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.reset()

    def reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.handles = [None] * len(self.buckets)

    def _on_grad(self, p):
        bi, off = self.where[p]
        self.flat[bi][off:off + p.numel()].copy_(p.grad.reshape(-1))  # cast to the communication dtype
        self.pending[bi] -= 1
        if self.pending[bi] == 0 and self.world > 1:
            self.handles[bi] = dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Waits for the bucket collectives and writes the averaged gradients back. Parameters that received no gradient
        in this step (their bucket never completed) are reduced here as zeros, so that all ranks stay in lock step."""
        for bi, b in enumerate(self.buckets):
            if self.pending[bi] != 0:
                for p in b:
                    if p.grad is None:
                        _, off = self.where[p]
                        self.flat[bi][off:off + p.numel()].zero_()
                if self.world > 1:
                    self.handles[bi] = dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        for bi, b in enumerate(self.buckets):
            if self.handles[bi] is not None:
                self.handles[bi].wait()
            if self.world > 1:
                for p in b:
                    _, off = self.where[p]
                    g = self.flat[bi][off:off + p.numel()].view_as(p).to(p.dtype) / self.world
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.copy_(g)
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()


def all_reduce_scalars(scalars, group=None):
    """Average a dict of 0-d tensors / floats over the ranks with ONE all-reduce (the reference: one per key, every iteration)."""
    keys = sorted(scalars)
    if not keys:
        return {}
    dev = next((v.device for v in scalars.values() if torch.is_tensor(v)), torch.device("cpu"))
    flat = torch.stack([torch.as_tensor(scalars[k], dtype=torch.float32, device=dev).reshape(()) for k in keys])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        flat = flat / dist.get_world_size(group)
    return {k: flat[i] for i, k in enumerate(keys)}
