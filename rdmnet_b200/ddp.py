"""Data-parallel training plumbing (BASELINE config 4): the reference wraps its model in
torch DistributedDataParallel (geotransformer/engine/base_trainer.py:181-191: fp32 gradients, 25 MB buckets) and
all-reduces every logged scalar separately each iteration (utils/torch.py:16-21, base_trainer.py:236). Here:

  BucketedGradAllReduce   parameter gradients are views into flat fp32 bucket buffers; as soon as a bucket's last gradient has been
                          accumulated (post-accumulate-grad hooks) the bucket is cast to bf16 (or sent as fp32) and its NCCL
                          all-reduce is launched asynchronously - it overlaps the rest of the backward pass -, and `finish()`
                          (before the optimizer step) waits, casts back and averages: two kernels per bucket, nothing per
                          parameter. 25.3 M parameters = 50.6 MB in bf16 per step.
  all_reduce_scalars      ONE collective for all logged scalars instead of one per key.

One process per GPU (torchrun); NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests. The hooks only pack,
cast and launch collectives: no arithmetic of the model runs here."""
import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    """Gradients live in flat fp32 bucket buffers: every parameter's `.grad` is a VIEW into its bucket (like torch DDP's
    gradient_as_bucket_view), so packing and unpacking cost nothing. A post-accumulate-grad hook counts the bucket down; when its
    last gradient has been accumulated the bucket is cast to the wire dtype (one kernel; skipped for fp32) and its all-reduce is
    launched asynchronously, overlapping the rest of the backward pass. `finish()` (before the optimizer step) waits, casts back and
    scales by 1/world: two kernels per bucket, none per parameter. Use `zero_grad()` of this object (one memset per bucket) instead
    of `optimizer.zero_grad(set_to_none=True)`, which would detach the views."""

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
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in self.buckets]
        self.wire = [f if comm_dtype == torch.float32 else torch.zeros_like(f, dtype=comm_dtype) for f in self.flat]
        self.where, self.views = {}, {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b:
                self.where[p] = bi
                self.views[p] = self.flat[bi][off:off + p.numel()].view_as(p)
                p.grad = self.views[p]
                off += p.numel()
        self.pending = [0] * len(self.buckets)
        self.handles = [None] * len(self.buckets)
        self.bytes_per_step = sum(w.numel() * w.element_size() for w in self.wire)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.reset()

    def reset(self):
        self.pending = [len(b) for b in self.buckets]
        self.handles = [None] * len(self.buckets)

    def zero_grad(self):
        for f in self.flat:
            f.zero_()
        for p, view in self.views.items():  # re-attach views that were dropped with set_to_none
            if p.grad is None:
                p.grad = view

    def _launch(self, bi):
        if self.world > 1:
            if self.wire[bi] is not self.flat[bi]:
                self.wire[bi].copy_(self.flat[bi])  # cast to the communication dtype
            self.handles[bi] = dist.all_reduce(self.wire[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, p):
        bi = self.where[p]
        view = self.views[p]
        if p.grad is not view and p.grad.data_ptr() != view.data_ptr():
            # somebody replaced the gradient tensor (optimizer.zero_grad(set_to_none=True), a manual `p.grad = None`): adopt the
            # freshly accumulated values and re-attach the view, instead of reducing a stale bucket slice
            view.copy_(p.grad)
            p.grad = view
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        """Waits for the bucket collectives and leaves the averaged gradients in place. A bucket whose countdown did not reach zero
        (a parameter received no gradient in this step: its slice is still zero) is reduced here, so that all ranks stay in lock
        step."""
        for bi, b in enumerate(self.buckets):
            if self.pending[bi] != 0:
                for p in b:  # no gradient this step and the old tensor dropped (set_to_none): contribute zeros, re-attach the view
                    if p.grad is None:
                        self.views[p].zero_()
                        p.grad = self.views[p]
                self._launch(bi)
        for bi in range(len(self.buckets)):
            if self.handles[bi] is not None:
                self.handles[bi].wait()
                if self.wire[bi] is not self.flat[bi]:
                    self.flat[bi].copy_(self.wire[bi])
                self.flat[bi].mul_(1.0 / self.world)
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
