"""Data-parallel plumbing: the only distributed step of the path is the gradient all-reduce
(reference: apex DistributedDataParallel(model, delay_allreduce=True) at train_tasks.py:497 — one flattened
all-reduce after backward, averaged over the world size; batch split per rank at task_utils.py:435-437).

Here the gradients already live in ONE flat fp32 buffer (engine.ParamStore.grad), so the all-reduce runs in place
on contiguous, fixed-address buckets (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


class FlatGradAllReducer:
    def __init__(self, flat_grad, n_buckets=8, group=None, align=1024):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = flat_grad.numel()
        step = max(align, (n + n_buckets - 1) // n_buckets)
        step = (step + align - 1) // align * align
        self.buckets = [flat_grad[i:min(i + step, n)] for i in range(0, n, step)]
        # NCCL has a fused average; gloo only sums
        self.use_avg = dist.is_initialized() and dist.get_backend(group) == "nccl"

    def allreduce(self, stream=None):
        """Averages the flat gradient buffer over all ranks, bucket by bucket (in place)."""
        if self.world == 1:
            return
        for b in self.buckets:
            if self.use_avg:
                dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
                b.div_(self.world)

    def allreduce_range(self, lo, hi, async_op=True):
        """Averages flat[lo:hi] over all ranks; returns the async work handle (or None for world 1). Used by the
        overlapped step: backward finishes the buffer from its end, ranges are reduced while backward continues."""
        if self.world == 1 or hi <= lo:
            return None
        t = self.flat[lo:hi]
        if self.use_avg:
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=False)
        t.div_(self.world)
        return w

    def allreduce_range_sync(self, lo, hi):
        """flat[lo:hi] averaged over all ranks, enqueued on the CURRENT stream (used under CUDA-graph capture)."""
        if self.world == 1 or hi <= lo:
            return
        t = self.flat[lo:hi]
        if self.use_avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def broadcast_params(self, flat_params, src=0):
        """Rank-`src` parameters to every rank (what apex DDP does at wrap time)."""
        if self.world > 1:
            dist.broadcast(flat_params, src=src, group=self.group)


class DistributedDataParallel:
    """Drop-in for `apex.parallel.DistributedDataParallel(model, delay_allreduce=True)` as the reference uses it
    (train_tasks.py:490-497): rank 0's parameters are broadcast at wrap time and every `loss.backward()` ends with ONE all-reduce
    (average over the world) of the flat fp32 gradient buffer. Not an nn.Module wrapper with hooks: the engine's backward calls
    the reducer itself. `.module` is the wrapped model, calls are forwarded."""

    def __init__(self, model, delay_allreduce=True, n_buckets=8, group=None):
        self.module = model
        eng = model.engine
        self.reducer = FlatGradAllReducer(eng.ps.grad, n_buckets=n_buckets, group=group)
        self.reducer.broadcast_params(eng.ps.flat)
        eng.shadow_clean = False
        model._ddp_reducer = self.reducer

    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self.module, name)


def shard_batch(global_batch, rank, world):
    """Per-rank batch like the reference: batch_size // world_size samples each (task_utils.py:435-437)."""
    per = global_batch // world
    return rank * per, per
