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

Two transports for the bucket all-reduce:
  * "nvlink" (default on CUDA when the process group spans one NVSwitch / NVLink box of <= 8 GPUs and torch's
    symmetric memory can map the peers): the library's own kernel (cvb_allreduce_mean_f32, csrc/allreduce.cu) over
    peer pointers. It uses no shared memory and few registers, so its CTAs run beside the convolution CTAs instead of
    waiting for -- and then delaying -- them as NCCL's kernels do; rank r reduces slice r in rank order and stores it to
    every rank: deterministic, bit-identical replicas. The flat gradient buffer then lives in symmetric memory and
    is reused step after step.
  * "nccl": torch.distributed.all_reduce (also the gloo path of the CPU tests). CVB_DP_BACKEND=nccl forces it.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

# Measurement knobs: CVB_BUCKET_MB sets the bucket size (a value larger than the gradient buffer =
# one all-reduce after the backward pass, nothing overlapped); CVB_DP_PAYLOAD=bf16 halves the bytes on the wire (the
# bucket is converted on the side stream; mean of bf16-rounded gradients); CVB_DP_PAYLOAD=none skips the exchange
# entirely (replicas drift apart: only for timing the compute without any communication).
_BUCKET_MB = float(os.environ["CVB_BUCKET_MB"]) if "CVB_BUCKET_MB" in os.environ else None  # None: per-transport default
_PAYLOAD = os.environ.get("CVB_DP_PAYLOAD", "fp32")


class GradReducer:
    """Bucketed mean all-reduce over a flat gradient buffer whose ranges become ready front to back."""

    DEFAULT_BUCKET_MB = 25.0

    def __init__(self, process_group=None, bucket_mb=None):
        if bucket_mb is None:
            bucket_mb = _BUCKET_MB if _BUCKET_MB is not None else self.DEFAULT_BUCKET_MB
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


class PeerReducer(GradReducer):
    """The same bucket stream, exchanged by the library's own NVLink kernel on a symmetric (peer-mapped) gradient
    buffer that this reducer owns. `buffer()` hands the plan the flat buffer of a backward pass."""

    # measured (tools/dp_ab.py, interleaved in one process, UNet 16 x 3 x 360 x 480 per GPU): small buckets on few CTAs
    # disturb the backward pass least -- see DESIGN.md section 6
    CTAS = int(os.environ.get("CVB_ALLREDUCE_CTAS", "32"))
    DEFAULT_BUCKET_MB = 8.0
    # CVB_NVLS=1: reduce inside the NVSwitch through the buffer's multicast address (multimem.ld_reduce / multimem.st)
    # where torch's symmetric memory provides one; 0 (default) = peer loads / stores
    NVLS = os.environ.get("CVB_NVLS", "0") != "0"

    def __init__(self, process_group=None, bucket_mb=None):
        super().__init__(process_group, bucket_mb)
        self.comm = None
        self.epoch = 0
        self.bucket = 0
        self.size = 0

    def _setup(self, size, device):
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        lib = _lib.load()
        if self.world > 8:
            raise RuntimeError("the NVLink reducer serves the (<= 8) GPUs of one box")
        with torch.cuda.device(device):
            self.flat_buf = symm.empty(size, dtype=torch.float32, device=device)
            self.flags = symm.empty(lib.cvb_comm_flag_words(), dtype=torch.int32, device=device)
            self.flat_buf.zero_()
            self.flags.zero_()
            h_buf = symm.rendezvous(self.flat_buf, group)
            h_flag = symm.rendezvous(self.flags, group)
            torch.cuda.synchronize(device)
            dist.barrier(group)  # every rank's pads are zero before anyone signals
        rank = dist.get_rank(group)
        self._bufs = (ctypes.c_void_p * self.world)(*[int(p) for p in h_buf.buffer_ptrs])
        self._flags = (ctypes.c_void_p * self.world)(*[int(p) for p in h_flag.buffer_ptrs])
        mc = int(getattr(h_buf, "multicast_ptr", 0) or 0) if self.NVLS else 0
        self.multicast = bool(mc)
        self.comm = _lib.Comm(ctypes.cast(self._bufs, ctypes.POINTER(ctypes.c_void_p)),
                              ctypes.cast(self._flags, ctypes.POINTER(ctypes.c_void_p)), ctypes.c_void_p(mc or None), rank,
                              self.world)
        self._handles = (h_buf, h_flag)
        self.size = size

    def buffer(self, size, device, params=()):
        """The flat gradient buffer of this backward pass: the symmetric buffer itself (conv-bias slices stay zero, every
        other range is overwritten by the pass). Gradients of an earlier pass that are still attached to parameters
        (gradient accumulation without zero_grad) alias it and are detached into private copies first."""
        if self.comm is None or self.size != size:
            self._setup(size, device)
        lo = self.flat_buf.data_ptr()
        hi = lo + self.flat_buf.numel() * 4
        for p in params:
            if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                p.grad = p.grad.clone()
        return self.flat_buf

    def begin(self, flat):
        super().begin(flat)
        self.epoch += 1
        self.bucket = 0

    def _launch(self, also_wait=None):
        if self.hi == self.lo:
            return
        lo, hi = self.lo, self.hi
        self.lo = self.hi
        self.buckets_launched += 1
        if _PAYLOAD == "none":
            return
        dev = self.flat.device
        self.side.wait_stream(torch.cuda.current_stream(dev))
        if also_wait is not None:
            self.side.wait_stream(also_wait)
        with torch.cuda.device(dev), torch.cuda.stream(self.side):
            rc = _lib.load().cvb_allreduce_mean_f32(ctypes.byref(self.comm), lo, hi - lo, self.bucket, self.epoch, self.CTAS,
                                                    ctypes.c_void_p(self.side.cuda_stream))
        if rc != 0:
            _lib.check(rc, "allreduce_mean_f32")
        self.bucket += 1

    def finish(self):
        self._launch()
        if self.flat is not None and self.bucket > 0:
            dev = self.flat.device
            cur = torch.cuda.current_stream(dev)
            cur.wait_stream(self.side)  # this rank's own kernels; then the arrival flags of everybody's slices
            with torch.cuda.device(dev):
                rc = _lib.load().cvb_allreduce_wait(ctypes.byref(self.comm), self.bucket, self.epoch,
                                                    ctypes.c_void_p(cur.cuda_stream))
            if rc != 0:
                _lib.check(rc, "allreduce_wait")
        self.works = []
        self.flat = None


def make_reducer(device, process_group=None, bucket_mb=None):
    """PeerReducer where it can work (CUDA, NCCL group of 2..8 ranks, symmetric memory importable), else GradReducer.
    CVB_DP_BACKEND=nccl / nvlink overrides (nvlink: fail loudly instead of falling back)."""
    want = os.environ.get("CVB_DP_BACKEND", "auto")
    ok = (dist.is_initialized() and device is not None and torch.device(device).type == "cuda"
          and dist.get_backend(process_group) == "nccl" and 2 <= dist.get_world_size(process_group) <= 8)
    if want == "nccl" or (want == "auto" and not ok):
        return GradReducer(process_group, bucket_mb)
    if not ok:
        raise RuntimeError("CVB_DP_BACKEND=nvlink needs an NCCL process group of 2..8 CUDA ranks on one box")
    try:
        import importlib
        importlib.import_module("torch.distributed._symmetric_memory")
    except Exception:
        if want == "nvlink":
            raise
        return GradReducer(process_group, bucket_mb)
    return PeerReducer(process_group, bucket_mb)


def data_parallel(module, process_group=None, bucket_mb=None, broadcast=True):
    """Marks a drop-in UNet / SegNet for data-parallel training and returns it (the module API is unchanged).

    broadcast=True copies rank 0's parameters and buffers to every rank first, so replicas start identical even when
    the processes were seeded differently.
    """
    device = next((p.device for p in module.parameters()), None)
    reducer = make_reducer(device, process_group, bucket_mb)
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
