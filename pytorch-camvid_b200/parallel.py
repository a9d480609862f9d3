"""Data-parallel training of the drop-in modules: one process per GPU, full weight replica, batch sharded by the
caller, bucketed gradient all-reduce (NCCL over NVLink / NVSwitch) overlapped with the rest of the backward pass.

Precedent in the reference: legacy/train_tpu.py:211-216,115 (torch_xla DataParallel over 8 cores, gradients
all-reduced in xm.optimizer_step, per-replica BatchNorm statistics). train.py itself is single-device.

How it plugs in: the execution plan (engine.Plan) writes every parameter gradient of a backward pass into ONE flat
fp32 buffer laid out in completion order (last layer first). As each block finishes, the plan reports the newly
complete range; GradReducer cuts the stream of ranges into buckets and launches one all-reduce per bucket on a side
stream while the compute stream carries on with the remaining dgrad / wgrad kernels. The compute stream waits for the
buckets only once, right before autograd hands the gradients to the optimizer.

BatchNorm uses per-replica batch statistics (what the reference's DP and torch DDP do). Gradients are averaged, which
equals the gradient of the global-mean loss when shards are equal-sized.
"""
import os

import torch
import torch.distributed as dist

# Measurement knobs: CVB_BUCKET_MB sets the bucket size (a value larger than the gradient buffer =
# one all-reduce after the backward pass, nothing overlapped); CVB_DP_PAYLOAD=bf16 halves the bytes on the wire (the
# bucket is converted on the side stream; mean of bf16-rounded gradients); CVB_DP_PAYLOAD=none skips the exchange
# entirely (replicas drift apart: only for timing the compute without any communication).
_BUCKET_MB = float(os.environ.get("CVB_BUCKET_MB", "25"))
_PAYLOAD = os.environ.get("CVB_DP_PAYLOAD", "fp32")


class GradReducer:
    """Bucketed mean all-reduce over a flat gradient buffer whose ranges become ready front to back."""

    def __init__(self, process_group=None, bucket_mb=None):
        bucket_mb = _BUCKET_MB if bucket_mb is None else bucket_mb
        if not dist.is_initialized():
            raise RuntimeError("camvid_b200.parallel: torch.distributed is not initialised")
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.backend = dist.get_backend(process_group)
        self.bucket_elems = max(1, int(bucket_mb * (1 << 20)) // 4)
        self.side = None
        self.flat = None
        self.lo = self.hi = 0
        self.works = []
        self.buckets_launched = 0  # of the last backward pass (introspection / tests)

    def begin(self, flat):
        self.flat, self.lo, self.hi, self.works = flat, 0, 0, []
        self.buckets_launched = 0
        if flat.is_cuda and self.side is None:
            self.side = torch.cuda.Stream(device=flat.device)

    def ready(self, lo, hi, also_wait=None):
        """flat[lo:hi] is final once the work enqueued so far on the current stream (and on `also_wait`, the plan's
        weight-gradient stream) has run. Ranges must arrive contiguously in increasing order."""
        if lo != self.hi:
            raise RuntimeError(f"camvid_b200.parallel: gradient range [{lo},{hi}) does not continue at {self.hi}")
        self.hi = hi
        if self.hi - self.lo >= self.bucket_elems:
            self._launch(also_wait)

    def _launch(self, also_wait=None):
        if self.hi == self.lo:
            return
        chunk = self.flat[self.lo:self.hi]
        self.lo = self.hi
        self.buckets_launched += 1
        if self.world == 1 or _PAYLOAD == "none":
            return
        avg = self.backend == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        if chunk.is_cuda:
            self.side.wait_stream(torch.cuda.current_stream(chunk.device))
            if also_wait is not None:
                self.side.wait_stream(also_wait)
            with torch.cuda.stream(self.side):
                if _PAYLOAD == "bf16":
                    half = chunk.to(torch.bfloat16)
                    work = dist.all_reduce(half, op=op, group=self.group, async_op=True)
                    work.wait()
                    chunk.copy_(half)
                else:
                    work = dist.all_reduce(chunk, op=op, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(chunk, op=op, group=self.group, async_op=True)
        self.works.append((work, chunk, avg))

    def finish(self):
        """Flush the tail bucket and make the current stream wait for every all-reduce."""
        self._launch()
        for work, chunk, avg in self.works:
            work.wait()
            if not avg:
                chunk.mul_(1.0 / self.world)
        if self.flat is not None and self.flat.is_cuda and self.side is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.side)
        self.works = []
        self.flat = None


def data_parallel(module, process_group=None, bucket_mb=None, broadcast=True):
    """Marks a drop-in UNet / SegNet for data-parallel training and returns it (the module API is unchanged).

    broadcast=True copies rank 0's parameters and buffers to every rank first, so replicas start identical even when
    the processes were seeded differently.
    """
    reducer = GradReducer(process_group, bucket_mb)
    if broadcast and reducer.world > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
    module.__dict__["_cvb_reducer"] = reducer
    return module


def all_reduce_confusion(cm, process_group=None):
    """Eval under data parallelism: sums a per-rank confusion matrix (int64 [C,C] tensor) over the ranks, in place."""
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=process_group)
    return cm
