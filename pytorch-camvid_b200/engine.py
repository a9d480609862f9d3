"""Execution plans behind the drop-in UNet / SegNet modules.

A plan owns every device buffer one (batch, height, width) configuration needs -- NHWC bf16 activations, raw conv
outputs, concat buffers written in place by their producers, pooling codes, per-layer BatchNorm vectors, packed bf16
weights, the wgrad split-K workspace -- and issues the kernel sequence of the forward and backward pass through the
C ABI (ops.py). Nothing is allocated per step except the flat fp32 gradient buffer handed to autograd.

Reference topology: models/unet.py:35-156 and models/segnet.py:19-119; one "block" is the reference's
BasicConv2d / BasicConv (conv3x3 pad 1 + BatchNorm2d + ReLU, models/unet.py:5-17, models/segnet.py:5-17).
"""
import collections
import itertools
import os
import weakref
from typing import List

import torch

from . import ops
from .ops import pad64


# Backward runs every weight-gradient kernel on a side stream: it only needs (x, dy) of its own block, while the critical
# path (dgrad -> BatchNorm backward of the previous block -> dgrad ...) alternates tensor-bound and HBM-bound kernels, so
# the wgrad MMAs overlap the memory-bound BatchNorm / pooling / upsampling passes. bench.py turns it off for its
# per-kernel timing pass (one stream = unambiguous event brackets).
OVERLAP_WGRAD = True
# Data-gradient kernels CAN emit the BatchNorm+ReLU backward statistics of the block they feed (cvb_conv_epilogue.bwd_*).
# Off by default: measured 1 % slower end to end -- the reduce pass it removes was already hidden under the side-stream
# weight gradient, while the heavier epilogue lengthens the data gradient on the critical path (CVB_FUSE_BWD=1 enables).
FUSE_BWD_STATS = os.environ.get("CVB_FUSE_BWD", "0") != "0"
# MaxPool backward emits the BatchNorm+ReLU backward reduction of the block it feeds (cvb_maxpool2x2_bwd_bn_reduce): the
# reduce pass over the largest activations of the encoder disappears (4 of 23 passes in UNet, 5 of 26 in SegNet).
FUSE_POOL_BWD_REDUCE = os.environ.get("CVB_FUSE_POOL_BWD", "1") != "0"
# The last block is fused with the module boundary: its BatchNorm+ReLU writes the fp32 NCHW logits directly
# (cvb_bn_relu_apply_nchw_f32) and the fp32 NCHW gradient of the loss is converted AND reduced for its BatchNorm backward
# in one pass (cvb_nchw_f32_to_nhwc_bf16_bn_reduce): two full-resolution HBM passes and two launches fewer per step.
FUSE_BOUNDARY = os.environ.get("CVB_FUSE_BOUNDARY", "1") != "0"
# UNet: the BatchNorm+ReLU of a block whose activation feeds ONLY a bilinear upsampling (bottleneck, up1-up3 second
# conv) is applied inside the upsampling kernel (cvb_bn_relu_bilinear2x_fwd); the activation itself is not written.
FUSE_UPSAMPLE_BN = os.environ.get("CVB_FUSE_UPSAMPLE_BN", "1") != "0"
# Visiting order of the BatchNorm passes (same results): each pass reads what the kernel before it just wrote front to
# back, and of a tensor larger than the 126 MB L2 only the END should still be cached, so a reversed pass could start on
# cached lines. Measured (tools/order_ab.py, profiles/r02x_order_ab_unet.txt): no gain in any combination (-0.02 ... +0.28
# ms on a 21.1 ms step, spread of the samples 0.4 ms) -- the L2 does not behave like an LRU stack under these streams.
# Off; the flag stays in the ABI for callers with other sizes.
REV_APPLY = os.environ.get("CVB_REV_APPLY", "0") != "0"
REV_BWD_REDUCE = os.environ.get("CVB_REV_BWD_REDUCE", "0") != "0"
REV_BWD_APPLY = os.environ.get("CVB_REV_BWD_APPLY", "0") != "0"
# Inspection aid (tests read every block's activation back from the plan): also write the activations that the fusions
# above make unnecessary.
MATERIALIZE_ACTIVATIONS = os.environ.get("CVB_MATERIALIZE_ACTIVATIONS", "0") != "0"
SERIALIZE_TENSOR_KERNELS = os.environ.get("CVB_SERIALIZE_TENSOR", "0") != "0"  # measured: see DESIGN.md knobs
# The side stream may trail the main stream by WGRAD_LAG blocks: the weight gradient of block k is enqueued after the
# data gradient of block k + LAG. The backward pass starts and ends with the full-resolution blocks, whose kernels on
# BOTH streams are HBM-bound (BatchNorm passes, cout = 64 convolutions, the 0.9 GB weight gradients), and spends its
# middle in the deep layers, where both streams are tensor-bound; a lag shifts the HBM-heavy weight gradients under the
# deep data gradients and the tensor-bound deep weight gradients under the HBM-bound passes of the last encoder blocks.
WGRAD_LAG = int(os.environ.get("CVB_WGRAD_LAG", "0"))


def narrow_channels(c):
    """Channels kept in memory for a layer of c outputs: 16-channel granularity (one MMA K step) up to the 64-wide GEMM
    chunk. TMA zero-fills the rest of the box, so 12 classes cost 16 channels of HBM traffic instead of 64."""
    return min(pad64(c), (c + 15) // 16 * 16)


class Block:
    """conv3x3(pad 1, bias) + BatchNorm2d + ReLU on NHWC bf16 views."""

    def __init__(self, plan, name, conv, bn, x, a, taps=9):
        self.plan, self.name, self.conv, self.bn = plan, name, conv, bn
        self.x, self.a, self.taps = x, a, taps
        self.cin, self.cout = conv.in_channels, conv.out_channels
        n, h, w, cin_mem = x.shape
        # channels of x in memory may stop short of the 64-wide GEMM chunk (the im2col'd first layer keeps 32): TMA
        # zero-fills the rest of the box; the packed weights are always padded to 64
        self.cin_pad = pad64(cin_mem)
        self.cout_pad = pad64(self.cout)  # GEMM padding; the activation buffers may hold fewer channels (12 -> 16)
        c_mem = a.shape[3]
        assert a.shape[:3] == x.shape[:3], (name, a.shape, x.shape)
        assert c_mem in (self.cout_pad, narrow_channels(self.cout)) and cin_mem % 16 == 0
        self.count = n * h * w
        dev = x.device
        self.y = torch.empty(n, h, w, c_mem, dtype=torch.bfloat16, device=dev)  # conv output, later dy
        self.vec = torch.zeros(4, self.cout_pad, device=dev)  # mean, invstd, scale, shift
        self.coef = torch.zeros(3, self.cout_pad, device=dev)
        self.wf = torch.empty(self.cout_pad, taps * self.cin_pad, dtype=torch.bfloat16, device=dev)
        self.wd = None if taps == 1 else torch.empty(self.cin_pad, 9 * self.cout_pad, dtype=torch.bfloat16, device=dev)
        self.wf_version = self.wd_version = None
        self.flops = 2.0 * 9 * self.cin * self.cout * self.count  # algorithmic FLOPs of one pass (un-padded channels)
        # Elementwise kernels only touch the channels that exist (rounded up to the 16-byte vector): the padded output
        # channels of y are exact zeros (zero weight rows), stay zero as dy, and are never read as activations. Matters
        # for the 12-class output layer, whose 64-channel-padded full-resolution tensors would otherwise cost 4x.
        self.ce = min(c_mem, (self.cout + 7) // 8 * 8)
        self.y_e, self.a_e = self.y[..., :self.ce], self.a[..., :self.ce]
        self.dy_k = self.y[..., :min(c_mem, narrow_channels(self.cout))]
        # the weight gradient reads dy through 16-channel granularity too when the layer is narrower than 64 channels
        self.dy_w = self.dy_k if (taps == 9 and self.cout_pad == 64) else self.y
        self.ws_bytes = ops.conv3x3_wgrad_workspace_bytes(x, self.dy_w, taps)
        self.c_ratio = self.cout / self.ce
        # offsets into the flat gradient buffer, assigned by the plan
        self.g_w = self.g_b = self.g_gamma = self.g_beta = None
        self._fuses = None  # whether this block's data gradient emits its consumer's backward statistics (lazy)

    # ---- parameters in reference order: conv.weight, conv.bias, bn.weight, bn.bias
    def params(self):
        return [self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias]

    def _weight_key(self):
        w = self.conv.weight
        return (w.data_ptr(), w._version)

    def _pack_f(self):
        key = self._weight_key()
        if key != self.wf_version:
            ops.pack_weights_fprop(self.conv.weight.detach(), self.taps, self.cout_pad, self.cin_pad, out=self.wf)
            self.wf_version = key

    def _pack_d(self):
        key = self._weight_key()
        if key != self.wd_version:
            ops.pack_weights_dgrad(self.conv.weight.detach(), self.cout_pad, self.cin_pad, out=self.wd)
            self.wd_version = key

    def fuses_boundary(self):
        """Whether this block, as the network's last one, can write the logits / read the loss gradient directly."""
        return FUSE_BOUNDARY and self.ce in (8, 16)

    def forward_train(self, pool_out=None, code=None, logits_out=None, defer_apply=False):
        """defer_apply: the consumer applies this block's BatchNorm+ReLU itself (from y, vec[2], vec[3])."""
        p, bn = self.plan, self.bn
        self._pack_f()
        parts = p.parts_view(self.cout_pad)
        ops.conv3x3(self.x, self.wf, self.y, taps=self.taps, stat_partials=parts, algo_flops=self.flops)
        ops.WORK_SCALE = self.c_ratio
        v = self.vec
        momentum = 0.1 if bn.momentum is None else bn.momentum
        track = bn.track_running_stats and bn.running_mean is not None
        ops.bn_finalize(parts, p.stat_rows, self.cout, self.cout_pad, self.count, bn.weight.detach(),
                        bn.bias.detach(), self.conv.bias.detach() if self.conv.bias is not None else None,
                        bn.running_mean if track else None, bn.running_var if track else None, momentum, bn.eps,
                        v[0], v[1], v[2], v[3])
        if logits_out is not None:
            # the network's last block: its activation is the fp32 NCHW tensor the module returns; the bf16 copy in
            # self.a is not written (backward takes the ReLU mask from y)
            ops.bn_relu_apply_nchw(self.y_e, v[2], v[3], logits_out)
            if MATERIALIZE_ACTIVATIONS:
                ops.bn_relu_apply(self.y_e, v[2], v[3], self.a_e)
        elif defer_apply and not MATERIALIZE_ACTIVATIONS:
            pass
        elif pool_out is not None:
            ops.bn_relu_maxpool2x2(self.y, v[2], v[3], self.a, pool_out, code)
        else:
            ops.bn_relu_apply(self.y_e, v[2], v[3], self.a_e, reverse=REV_APPLY)
        ops.WORK_SCALE = 1.0

    def forward_eval(self):
        """BatchNorm folded into the conv epilogue with the running statistics (models/unet.py:12 in eval mode)."""
        bn = self.bn
        self._pack_f()
        key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               self.conv.bias._version if self.conv.bias is not None else 0, bn.weight.data_ptr())
        if getattr(self, "_fold_key", None) != key:
            with torch.no_grad():
                scale = torch.zeros(self.cout_pad, device=self.x.device)
                shift = torch.zeros(self.cout_pad, device=self.x.device)
                s = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
                b = self.conv.bias.detach() if self.conv.bias is not None else 0.0
                scale[:self.cout] = s
                shift[:self.cout] = bn.bias.detach() + (b - bn.running_mean) * s
            self._fold = (scale, shift)
            self._fold_key = key
        ops.conv3x3(self.x, self.wf, self.a, taps=self.taps, scale=self._fold[0], shift=self._fold[1], relu=True,
                    algo_flops=self.flops)

    def backward(self, da, dx, flat, consumer=None, stats_ready=False):
        """da: gradient w.r.t. self.a (same view geometry); dx: view receiving the gradient w.r.t. self.x or None.
        consumer: the block whose activation IS self.x (dx is its `da`, nothing else accumulates into it): where the
        data-gradient kernel can, it emits that block's BatchNorm+ReLU backward reduction from its epilogue. Returns True
        if it did; the caller passes that as `stats_ready` to the consumer's backward, which then skips its reduce pass."""
        p, v = self.plan, self.vec
        parts = p.parts_view(self.ce)  # the reduce kernel lays its partial rows out with the view's channel count
        ops.WORK_SCALE = self.c_ratio
        if self.ce != da.shape[3]:
            da = da[..., :self.ce]
        # stats_ready: False = reduce here; True = the data-gradient conv that produced da emitted the partial sums
        # (stat_rows of them); an int = another producer did (that many rows)
        rows = p.stat_rows if stats_ready is True else stats_ready
        if stats_ready is False:
            ops.bn_relu_bwd_reduce(da, self.y_e, v[2], v[3], parts, p.reduce_rows, reverse=REV_BWD_REDUCE)
            rows = p.reduce_rows
        dgamma = flat[self.g_gamma:self.g_gamma + self.cout]
        dbeta = flat[self.g_beta:self.g_beta + self.cout]
        ops.bn_bwd_finalize(parts, rows, self.cout, self.ce, self.count, self.bn.weight.detach(), v[0],
                            v[1], dgamma, dbeta, self.coef)
        ops.bn_relu_bwd_apply(da, self.y_e, v[2], v[3], self.coef, self.y_e, reverse=REV_BWD_APPLY)  # y now holds dy
        ops.WORK_SCALE = 1.0
        dw = flat[self.g_w:self.g_w + self.conv.weight.numel()].view_as(self.conv.weight)
        # (conv bias feeds a batch-stat BatchNorm: its gradient is exactly 0 -- the flat buffer starts zeroed)
        if p.wstream is None:
            ops.conv3x3_wgrad(self.x, self.dy_w, dw, taps=self.taps, workspace=p.workspace, algo_flops=self.flops)
        fused = False
        if dx is not None:
            self._pack_d()
            if p.wstream is not None and SERIALIZE_TENSOR_KERNELS:
                # experiment: the data gradient waits for the previous block's weight gradient instead of sharing the
                # SMs with its tail
                torch.cuda.current_stream(p.device).wait_stream(p.wstream)
            if consumer is not None and FUSE_BWD_STATS:
                if self._fuses is None:
                    self._fuses = (consumer.ce == consumer.cout_pad == dx.shape[3]
                                   and ops.conv3x3_fuses_bwd_stats(self.dy_k, dx))
                fused = self._fuses
            # dy of a padded layer (12 -> 64 output channels) is read through its real channels only: the data gradient
            # then skips the all-zero K steps (self.dy_k is a 16-channel-granular slice of y)
            if fused:
                ops.conv3x3(self.dy_k, self.wd, dx, algo_flops=self.flops,
                            bwd=(consumer.y_e, consumer.vec[2], consumer.vec[3], p.parts_view(consumer.ce)))
            else:
                ops.conv3x3(self.dy_k, self.wd, dx, algo_flops=self.flops)
        if p.wstream is not None:
            # The side stream picks the weight gradient up AFTER the data gradient: started together the two tensor-bound
            # kernels only split the SMs between them (measured: same finish time as back to back) and the HBM-bound
            # BatchNorm passes of the next block then run alone. Started here, the weight gradient's MMAs run under
            # those passes (they need no shared memory and co-reside with the wgrad CTAs). With WGRAD_LAG > 0 it is
            # queued instead and enqueued LAG blocks later (Plan.drain_wgrads).
            p.pending_wgrads.append((self, dw))
            p.drain_wgrads(WGRAD_LAG)
        return fused

    def launch_wgrad(self, dw):
        p = self.plan
        ops.conv3x3_wgrad(self.x, self.dy_w, dw, taps=self.taps, workspace=p.workspace, algo_flops=self.flops)


class Plan:
    """Buffers + kernel sequences for one input geometry. Subclasses define the topology."""

    def __init__(self, module, n, h, w, device):
        self.module, self.n, self.h, self.w, self.device = module, n, h, w, device
        self.stat_rows = ops.stat_rows()
        self.reduce_rows = int(os.environ.get("CVB_REDUCE_ROWS_PER_SM", "4")) * ops.sm_count()  # grid of the BN-backward reduction
        self.parts = torch.empty(max(self.stat_rows, self.reduce_rows), 2, 1024, device=device)
        self.blocks = []  # in forward order
        self.workspace = None
        self.generation = 0           # bumped by every forward: the plan's activation buffers are overwritten
        self.busy_generation = None   # generation whose saved activations a pending backward still needs
        self.done_generation = None   # generation whose backward has run (y buffers now hold dy: no second backward)
        self.generation_time = 0      # when the plan was last handed to a forward (pool replacement order)
        self.reducer = None  # parallel.GradReducer of the module during a backward pass (data parallelism)
        self.wstream = None  # side stream of the weight-gradient kernels during a backward pass
        self._wstream = None
        self.pending_wgrads = collections.deque()  # (block, dw) whose weight gradient is not enqueued yet (WGRAD_LAG)
        self.handle = next(_HANDLES)  # how the dispatcher ops below address this plan
        _PLANS[self.handle] = self

    def input_buffer(self, cin):
        """NHWC bf16 operand of the first convolution and its tap count. Up to 7 input channels (CamVid: 3) the 3x3
        neighbourhood is gathered into the channel dimension (im2col: K = 9 * cin <= 64 becomes ONE GEMM tap); wider
        inputs take the ordinary 9-tap path on a 64-channel-padded copy."""
        self.first_taps = 1 if 9 * cin <= 64 else 9
        c_mem = min(64, (9 * cin + 15) // 16 * 16) if self.first_taps == 1 else pad64(cin)
        return self.buf(self.h, self.w, c_mem)

    def load_input(self, x):
        if self.first_taps == 1:
            ops.im2col3x3(x, self.cols)
        else:
            ops.nchw_to_nhwc(x, self.cols)

    def drain_wgrads(self, keep):
        """Enqueues the queued weight gradients on the side stream, oldest first, until at most `keep` are left; each
        starts after everything the main stream has been given so far. A block's gradient range is reported to the
        data-parallel reducer when its weight gradient is enqueued (the ranges stay in layout order: FIFO)."""
        while len(self.pending_wgrads) > keep:
            b, dw = self.pending_wgrads.popleft()
            ready = torch.cuda.Event()
            ready.record()
            self.wstream.wait_event(ready)
            with torch.cuda.stream(self.wstream):
                b.launch_wgrad(dw)
            if self.reducer is not None:
                self.reducer.ready(b.g_w, b.g_end, also_wait=self.wstream)

    def parts_view(self, c):
        rows = self.parts.shape[0]
        return self.parts.view(-1)[:rows * 2 * c].view(rows, 2, c)

    def buf(self, h, w, c, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(self.n, h, w, c, dtype=torch.bfloat16, device=self.device)

    def add(self, name, seq, x, a, taps=9):
        conv, bn = seq[0], seq[1]
        b = Block(self, name, conv, bn, x, a, taps)
        self.blocks.append(b)
        return b

    def finish(self):
        self.workspace = torch.empty(max(b.ws_bytes for b in self.blocks), dtype=torch.uint8, device=self.device)
        # flat gradient layout in backward completion order (last block first): contiguous ready ranges for the
        # data-parallel all-reduce buckets
        off = 0
        for b in reversed(self.blocks):
            b.g_w = off
            off += b.conv.weight.numel()
            if b.conv.bias is not None:
                b.g_b = off
                off += b.cout
            b.g_gamma = off
            off += b.cout
            b.g_beta = off
            off += b.cout
            b.g_end = off
        self.flat_size = off

    def pack_weights(self):
        """Refreshes the bf16 GEMM operands of every block whose fp32 weight changed (the optimizer bumps the version
        every step): all 3x3 layers in one launch, the im2col'd first layer separately."""
        stale = [b for b in self.blocks if b._weight_key() != b.wf_version]
        if torch.cuda.is_current_stream_capturing():
            stale = list(self.blocks)  # a captured step re-packs on every replay: its optimizer moves the weights
        if not stale:
            return
        batch = [b for b in stale if b.taps == 9]
        if batch:
            ptrs = tuple(b.conv.weight.data_ptr() for b in batch)
            if getattr(self, "_pack_key", None) != ptrs:
                self._pack_tab = ops.pack_table([(b.conv.weight.detach(), b.wf, b.wd, b.cout_pad, b.cin_pad)
                                                 for b in batch], self.device)
                self._pack_key = ptrs
                self._pack_dims = (max(b.cout_pad for b in batch), max(b.cin_pad for b in batch),
                                   sum(b.conv.weight.numel() * 4 + b.wf.numel() * 2 + b.wd.numel() * 2 for b in batch))
            ops.pack_weights_batch(self._pack_tab, *self._pack_dims)
            for b in batch:
                b.wf_version = b.wd_version = b._weight_key()
        for b in stale:
            if b.taps != 9:
                b._pack_f()

    def _logits_train(self, block):
        """Train-mode forward of the last block + the fp32 NCHW logits handed back to autograd."""
        logits = torch.empty(self.n, self.class_num, self.h, self.w, device=self.device)
        if block.fuses_boundary():
            block.forward_train(logits_out=logits)
        else:
            block.forward_train()
            ops.nhwc_to_nchw(self.out_a, logits)
        return logits

    def _enter_backward(self, block, dlogits):
        """fp32 NCHW gradient of the loss -> bf16 NHWC `da` of the last block. Returns the `stats_ready` argument of that
        block's backward: the number of partial rows when the conversion also emitted its BatchNorm backward reduction."""
        da = self.d_out_a[..., :block.ce]
        if block.fuses_boundary():
            ops.nchw_to_nhwc_bn_reduce(dlogits, da, block.y_e, block.vec[2], block.vec[3], self.parts_view(block.ce),
                                       self.reduce_rows)
            return self.reduce_rows
        ops.nchw_to_nhwc(dlogits, da)
        return False

    def param_list(self):
        out = []
        for b in self.blocks:
            out += [q for q in b.params() if q is not None]
        return out

    def grads_for(self, flat):
        out = []
        for b in self.blocks:
            out.append(flat[b.g_w:b.g_w + b.conv.weight.numel()].view_as(b.conv.weight))
            if b.conv.bias is not None:
                out.append(flat[b.g_b:b.g_b + b.cout])
            out.append(flat[b.g_gamma:b.g_gamma + b.cout])
            out.append(flat[b.g_beta:b.g_beta + b.cout])
        return out

    def _bump_batches_tracked(self):
        t = [b.bn.num_batches_tracked for b in self.blocks if b.bn.num_batches_tracked is not None]
        if t:
            torch._foreach_add_(t, 1)
        # bn_finalize updated the running statistics through raw pointers: bump their versions like an in-place op
        # would (the eval-mode fold of BatchNorm into the conv epilogue is cached on them)
        stats = [s_ for b in self.blocks for s_ in (b.bn.running_mean, b.bn.running_var) if s_ is not None]
        if stats:
            torch.autograd.graph.increment_version(stats)

    def _begin_backward(self):
        """Flat fp32 gradient buffer of this pass; under data parallelism its ranges are all-reduced as they fill."""
        self.reducer = self.module.__dict__.get("_cvb_reducer")
        if self.reducer is not None and hasattr(self.reducer, "buffer"):
            # NVLink reducer: the pass writes straight into the peer-mapped buffer the all-reduce kernel works on
            flat = self.reducer.buffer(self.flat_size, self.device, self.param_list())
        else:
            flat = torch.zeros(self.flat_size, device=self.device)  # one memset instead of a fill per conv-bias slice
        if self.reducer is not None:
            self.reducer.begin(flat)
        self.pending_wgrads.clear()
        if OVERLAP_WGRAD:
            if self._wstream is None:
                self._wstream = torch.cuda.Stream(device=self.device)
            self.wstream = self._wstream
            self.wstream.wait_stream(torch.cuda.current_stream(self.device))  # flat is allocated, workspace is free
        else:
            self.wstream = None
        return flat

    def _done(self, b):
        # with the side stream on, drain_wgrads reports the range when the block's weight gradient is enqueued
        if self.reducer is not None and self.wstream is None:
            self.reducer.ready(b.g_w, b.g_end, also_wait=None)

    def _end_backward(self, flat):
        if self.wstream is not None:
            self.drain_wgrads(0)
            torch.cuda.current_stream(self.device).wait_stream(self.wstream)
            self.wstream = None
        if self.reducer is not None:
            self.reducer.finish()
            self.reducer = None
        return flat


class UNetPlan(Plan):
    """models/unet.py:94-156. Skip concatenation is by construction: the encoder block and the up-conv block write
    their activations straight into channel slices of the shared concat buffer ([up | skip], upsampled branch first,
    models/unet.py:124); F.pad (models/unet.py:120-123) is the zero border of that buffer."""

    def __init__(self, m, n, h, w, device):
        super().__init__(m, n, h, w, device)
        self.class_num = m.class_num
        hs = [h]
        ws = [w]
        for _ in range(4):
            hs.append(hs[-1] // 2)
            ws.append(ws[-1] // 2)
        ch = [64, 128, 256, 512, 1024]
        self.cols = self.input_buffer(m.input_channels)  # CamVid: im2col of the input, 27 real channels of 32
        # concat buffers at levels 0..3: channels [0, ch[l]) = upsampled branch, [ch[l], 2 ch[l]) = encoder skip
        self.cat = [self.buf(hs[l], ws[l], 2 * ch[l], zero=True) for l in range(4)]
        self.dcat = [self.buf(hs[l], ws[l], 2 * ch[l]) for l in range(4)]
        downs = [m.down1, m.down2, m.down3, m.down4, m.down5]
        self.enc = []
        x = self.cols
        self.pooled, self.dpooled, self.enc_mid, self.d_enc_mid, self.codes = [], [], [], [], []
        for l in range(5):
            mid = self.buf(hs[l], ws[l], ch[l])
            self.enc_mid.append(mid)
            self.d_enc_mid.append(torch.empty_like(mid))
            b0 = self.add(f"down{l + 1}.0", downs[l][0].conv, x, mid, taps=self.first_taps if l == 0 else 9)
            if l < 4:
                out = self.cat[l][..., ch[l]:]
                pooled = self.buf(hs[l + 1], ws[l + 1], ch[l])
                self.pooled.append(pooled)
                self.dpooled.append(torch.empty_like(pooled))
                # window codes (1 byte per pooled element): the fused pool-backward + BatchNorm-reduce kernel reads them
                self.codes.append(torch.empty(pooled.shape, dtype=torch.uint8, device=device) if FUSE_POOL_BWD_REDUCE else None)
            else:
                out = self.buf(hs[l], ws[l], ch[l])
                self.bott, self.dbott = out, torch.empty_like(out)
            b1 = self.add(f"down{l + 1}.1", downs[l][1].conv, mid, out)
            self.enc.append((b0, b1))
            if l < 4:
                x = pooled
        ups = [(m.upsample1, m.up1), (m.upsample2, m.up2), (m.upsample3, m.up3), (m.upsample4, m.up4)]
        self.dec = []
        x = self.bott
        dx = self.dbott
        for i, (upm, seq) in enumerate(ups):
            l = 3 - i  # level of the concat buffer this stage writes into
            hu, wu = 2 * x.shape[1], 2 * x.shape[2]
            c_in = x.shape[3]
            up = self.buf(hu, wu, c_in)
            dup = torch.empty_like(up)
            dh, dw = hs[l] - hu, ws[l] - wu
            top, left = dh // 2, dw // 2  # F.pad offsets, models/unet.py:122-123
            win = (slice(None), slice(top, top + hu), slice(left, left + wu), slice(0, ch[l]))
            bu = self.add(f"upsample{i + 1}", upm.conv.conv, up, self.cat[l][win])
            m0 = self.buf(hs[l], ws[l], ch[l])
            m1 = self.buf(hs[l], ws[l], ch[l])
            b0 = self.add(f"up{i + 1}.0", seq[0].conv, self.cat[l], m0)
            b1 = self.add(f"up{i + 1}.1", seq[1].conv, m0, m1)
            self.dec.append(dict(src=x, dsrc=dx, up=up, dup=dup, win=win, bu=bu, b0=b0, b1=b1, m0=m0, m1=m1,
                                 dm0=torch.empty_like(m0), dm1=torch.empty_like(m1), level=l))
            x, dx = m1, self.dec[-1]["dm1"]
        # 12 classes: 16 channels in memory where the transposed cout = 64 kernel runs (even height), else 64
        out_c = narrow_channels(self.class_num) if (h % 2 == 0 and w >= 8) else pad64(self.class_num)
        self.out_a = self.buf(h, w, out_c)
        self.d_out_a = torch.empty_like(self.out_a)
        self.b_out = self.add("output", m.output.conv, x, self.out_a)
        self.finish()

    def forward(self, x, train):
        self.pack_weights()
        self.load_input(x)
        fuse_up = train and FUSE_UPSAMPLE_BN
        for l, (b0, b1) in enumerate(self.enc):
            if train:
                b0.forward_train()
                if l < 4:
                    b1.forward_train(pool_out=self.pooled[l], code=self.codes[l])
                else:
                    b1.forward_train(defer_apply=fuse_up)  # the bottleneck feeds only the first upsampling
            else:
                b0.forward_eval()
                b1.forward_eval()
                if l < 4:
                    ops.maxpool2x2(b1.a, self.pooled[l])
        prev = self.enc[4][1]  # the block whose activation is upsampled
        for i, d in enumerate(self.dec):
            if fuse_up:
                ops.bn_relu_bilinear2x(prev.y_e, prev.vec[2], prev.vec[3], d["up"])
            else:
                ops.bilinear2x(d["src"], d["up"])
            for b in (d["bu"], d["b0"], d["b1"]):
                if not train:
                    b.forward_eval()
                else:
                    # up1-up3: the second conv's activation feeds only the next upsampling
                    b.forward_train(defer_apply=fuse_up and b is d["b1"] and i < len(self.dec) - 1)
            prev = d["b1"]
        if train:
            logits = self._logits_train(self.b_out)
            self._bump_batches_tracked()
            return logits
        self.b_out.forward_eval()
        logits = torch.empty(self.n, self.class_num, self.h, self.w, device=self.device)
        ops.nhwc_to_nchw(self.out_a, logits)
        return logits

    def backward(self, dlogits):
        flat = self._begin_backward()
        ready = self._enter_backward(self.b_out, dlogits)
        last = self.dec[-1]
        ready = self.b_out.backward(self.d_out_a, last["dm1"], flat, consumer=last["b1"], stats_ready=ready)
        self._done(self.b_out)
        for d in reversed(self.dec):
            l = d["level"]
            ready = d["b1"].backward(d["dm1"], d["dm0"], flat, consumer=d["b0"], stats_ready=ready)
            self._done(d["b1"])
            d["b0"].backward(d["dm0"], self.dcat[l], flat, stats_ready=ready)  # dcat feeds two blocks: not fused
            self._done(d["b0"])
            d["bu"].backward(self.dcat[l][d["win"]], d["dup"], flat)
            self._done(d["bu"])
            ops.bilinear2x_bwd(d["dup"], d["dsrc"])
            ready = False  # the next stage's dm1 comes out of the upsampling backward
        for l in range(4, -1, -1):
            b0, b1 = self.enc[l]
            if l == 4:
                da = self.dbott
            else:
                ch = self.cat[l].shape[3] // 2
                da = self.dcat[l][..., ch:]
                # encoder activation feeds both the skip (already in dcat) and the pool: add the pool path
                if self.codes[l] is not None and b1.ce == da.shape[3]:
                    ops.maxpool2x2_bwd_bn_reduce(self.dpooled[l], da, b1.y_e, b1.vec[2], b1.vec[3],
                                                 self.parts_view(b1.ce), self.reduce_rows, code=self.codes[l],
                                                 accumulate=True)
                    pooled_ready = self.reduce_rows
                else:
                    ops.maxpool2x2_bwd(self.dpooled[l], da, x=b1.a, accumulate=True)
                    pooled_ready = False
            ready = b1.backward(da, self.d_enc_mid[l], flat, consumer=b0, stats_ready=pooled_ready if l < 4 else False)
            self._done(b1)
            b0.backward(self.d_enc_mid[l], self.dpooled[l - 1] if l > 0 else None, flat, stats_ready=ready)
            self._done(b0)
        return self._end_backward(flat)


class SegNetPlan(Plan):
    """models/segnet.py:82-119. MaxPool indices are 1-byte window codes local to the plan (the reference never
    exposes idx1..idx5 either); MaxUnpool writes into a plane of the saved encoder shape (output_size=fmK)."""

    ENC = [2, 2, 3, 3, 3]

    def __init__(self, m, n, h, w, device):
        super().__init__(m, n, h, w, device)
        self.class_num = m.class_num
        encs = [m.encoder1, m.encoder2, m.encoder3, m.encoder4, m.encoder5]
        decs = [m.decoder5, m.decoder4, m.decoder3, m.decoder2, m.decoder1]
        self.cols = self.input_buffer(m.input_channels)  # CamVid: im2col of the input, 27 real channels of 32
        self.stages = []  # encoder stages: blocks, activations, pooled, code
        x = self.cols
        ch_, cw_ = h, w
        for s, seq in enumerate(encs):
            blocks, acts, dacts = [], [], []
            for j, bc in enumerate(seq):
                a = self.buf(ch_, cw_, pad64(bc.conv.out_channels))
                blocks.append(self.add(f"encoder{s + 1}.{j}", (bc.conv, bc.bn), x, a,
                                       taps=self.first_taps if (s == 0 and j == 0) else 9))
                acts.append(a)
                dacts.append(torch.empty_like(a))
                x = a
            pooled = self.buf(ch_ // 2, cw_ // 2, x.shape[3])
            code = torch.empty(n, ch_ // 2, cw_ // 2, x.shape[3], dtype=torch.uint8, device=device)
            self.stages.append(dict(blocks=blocks, acts=acts, dacts=dacts, pooled=pooled, dpooled=torch.empty_like(pooled),
                                    code=code, shape=(ch_, cw_)))
            x = pooled
            ch_, cw_ = ch_ // 2, cw_ // 2
        self.dstages = []
        dx = self.stages[-1]["dpooled"]
        for i, seq in enumerate(decs):
            st = self.stages[4 - i]
            hh, ww = st["shape"]
            un = self.buf(hh, ww, x.shape[3])
            dun = torch.empty_like(un)
            blocks, acts, dacts = [], [], []
            xin = un
            for j, bc in enumerate(seq):
                last = i == len(decs) - 1 and j == len(seq) - 1  # the class logits: 16 channels in memory (see UNetPlan)
                narrow = last and hh % 2 == 0 and ww >= 8
                a = self.buf(hh, ww, narrow_channels(bc.conv.out_channels) if narrow else pad64(bc.conv.out_channels))
                blocks.append(self.add(f"decoder{5 - i}.{j}", (bc.conv, bc.bn), xin, a))
                acts.append(a)
                dacts.append(torch.empty_like(a))
                xin = a
            self.dstages.append(dict(src=x, dsrc=dx, un=un, dun=dun, code=st["code"], blocks=blocks, acts=acts,
                                     dacts=dacts))
            x, dx = xin, dacts[-1]
        self.out_a, self.d_out_a = x, dx
        self.finish()

    def forward(self, x, train):
        self.pack_weights()
        self.load_input(x)
        for st in self.stages:
            bl = st["blocks"]
            for j, b in enumerate(bl):
                if train:
                    if j == len(bl) - 1:
                        b.forward_train(pool_out=st["pooled"], code=st["code"])
                    else:
                        b.forward_train()
                else:
                    b.forward_eval()
            if not train:
                ops.maxpool2x2(bl[-1].a, st["pooled"], st["code"])
        last = self.dstages[-1]["blocks"][-1]
        logits = None
        for ds in self.dstages:
            ops.maxunpool2x2(ds["src"], ds["code"], ds["un"])
            for b in ds["blocks"]:
                if not train:
                    b.forward_eval()
                elif b is last:
                    logits = self._logits_train(b)
                else:
                    b.forward_train()
        if train:
            self._bump_batches_tracked()
            return logits
        logits = torch.empty(self.n, self.class_num, self.h, self.w, device=self.device)
        ops.nhwc_to_nchw(self.out_a, logits)
        return logits

    def backward(self, dlogits):
        flat = self._begin_backward()
        entry = self._enter_backward(self.dstages[-1]["blocks"][-1], dlogits)
        for ds in reversed(self.dstages):
            bl = ds["blocks"]
            ready = entry if ds is self.dstages[-1] else False
            for j in range(len(bl) - 1, -1, -1):
                ready = bl[j].backward(ds["dacts"][j], ds["dacts"][j - 1] if j > 0 else ds["dun"], flat,
                                       consumer=bl[j - 1] if j > 0 else None, stats_ready=ready)
                self._done(bl[j])
            ops.maxunpool2x2_bwd(ds["dun"], ds["code"], ds["dsrc"])
        for s in range(4, -1, -1):
            st = self.stages[s]
            bl = st["blocks"]
            if FUSE_POOL_BWD_REDUCE and bl[-1].ce == st["dacts"][-1].shape[3]:
                ops.maxpool2x2_bwd_bn_reduce(st["dpooled"], st["dacts"][-1], bl[-1].y_e, bl[-1].vec[2], bl[-1].vec[3],
                                             self.parts_view(bl[-1].ce), self.reduce_rows, code=st["code"])
                ready = self.reduce_rows
            else:
                ops.maxpool2x2_bwd(st["dpooled"], st["dacts"][-1], code=st["code"])
                ready = False
            for j in range(len(bl) - 1, -1, -1):
                if j > 0:
                    dx = st["dacts"][j - 1]
                else:
                    dx = self.stages[s - 1]["dpooled"] if s > 0 else None
                ready = bl[j].backward(st["dacts"][j], dx, flat, consumer=bl[j - 1] if j > 0 else None,
                                       stats_ready=ready)
                self._done(bl[j])
        return self._end_backward(flat)


class BlockPlan(Plan):
    """One reference layer on its own -- BasicConv2d (models/unet.py:5-17), BasicConv (models/segnet.py:5-17) or
    UpSample2d (bilinear x2 + BasicConv2d, models/unet.py:19-32) -- for callers that use a submodule of the drop-in
    network directly (`net.down1(x)`, feature extraction, a hand-written forward). fp32 NCHW in / out like the reference
    layer, the same kernels as inside the whole-network plans, gradients w.r.t. the input and the four parameters."""

    def __init__(self, layer, conv, bn, upsample, n, h, w, device):
        super().__init__(layer, n, h, w, device)
        self.class_num = conv.out_channels
        self.upsample = upsample
        cin = conv.in_channels
        self.src = self.buf(h, w, pad64(cin))
        self.dsrc = torch.empty_like(self.src)
        if upsample:
            self.up, self.dup = self.buf(2 * h, 2 * w, pad64(cin)), self.buf(2 * h, 2 * w, pad64(cin))
            xin, self.dxin, (ho, wo) = self.up, self.dup, (2 * h, 2 * w)
        else:
            xin, self.dxin, (ho, wo) = self.src, self.dsrc, (h, w)
        self.ho, self.wo = ho, wo
        self.out_a = self.buf(ho, wo, pad64(conv.out_channels))
        self.d_out_a = torch.empty_like(self.out_a)
        self.block = self.add("block", (conv, bn), xin, self.out_a)
        self.finish()

    def forward(self, x, train):
        self.pack_weights()
        ops.nchw_to_nhwc(x, self.src)
        if self.upsample:
            ops.bilinear2x(self.src, self.up)
        if train:
            self.block.forward_train()
            self._bump_batches_tracked()
        else:
            self.block.forward_eval()
        out = torch.empty(self.n, self.class_num, self.ho, self.wo, device=self.device)
        ops.nhwc_to_nchw(self.out_a, out)
        return out

    def backward(self, dout, need_dx=True):
        flat = self._begin_backward()
        ops.nchw_to_nhwc(dout, self.d_out_a)
        self.block.backward(self.d_out_a, self.dxin if need_dx else None, flat)
        self._done(self.block)
        dx = None
        if need_dx:
            if self.upsample:
                ops.bilinear2x_bwd(self.dup, self.dsrc)
            dx = torch.empty(self.n, self.block.cin, self.h, self.w, device=self.device)
            ops.nhwc_to_nchw(self.dsrc, dx)
        return self._end_backward(flat), dx


# ---------------------------------------------------------------------------------------------------------------------
# torch custom-op layer. The whole network is ONE dispatcher op per direction (`camvid_b200::net_forward` /
# `camvid_b200::net_backward`, registered through torch.library with a fake (shape-only) implementation and an autograd
# formula), so the drop-in modules are ordinary citizens of autograd AND of torch.jit.trace -- utils.visualize_network's
# `writer.add_graph(net, tensor)` (utils.py:10-13, train.py:97-98) records one opaque node instead of failing. Plans are
# addressed by an integer handle because dispatcher arguments are tensors and scalars.
_PLANS = weakref.WeakValueDictionary()  # handle -> plan (the owning module keeps the plan alive)
_HANDLES = itertools.count(1)


def _plan_of(handle):
    plan = _PLANS.get(handle)
    if plan is None:
        raise RuntimeError(f"camvid_b200: execution plan {handle} no longer exists (its module was freed)")
    return plan


@torch.library.custom_op("camvid_b200::net_forward", mutates_args=())
def net_forward_op(x: torch.Tensor, params: List[torch.Tensor], plan_handle: int, train: bool) -> torch.Tensor:
    """x fp32 NCHW -> logits fp32 NCHW. `params` (reference order: conv.weight, conv.bias, bn.weight, bn.bias per block)
    are listed so that autograd routes their gradients; the kernels read them through the plan. Train mode also updates
    the BatchNorm running statistics of the owning module, exactly like nn.BatchNorm2d."""
    plan = _plan_of(plan_handle)
    plan.generation += 1  # every forward overwrites the plan's activation buffers
    with ops.on_device(x):  # kernels go to the CURRENT device's stream: make the tensor's device current
        return plan.forward(x, train)


@net_forward_op.register_fake
def _net_forward_fake(x, params, plan_handle, train):
    return x.new_empty(x.shape[0], _plan_of(plan_handle).class_num, x.shape[2], x.shape[3])


def _check_generation(plan, generation):
    if generation != plan.generation:
        raise RuntimeError("camvid_b200: backward() of a forward pass whose saved activations were overwritten by a "
                           "later forward of the same module and input shape (each shape keeps CVB_PLANS_PER_SHAPE = "
                           f"{PLANS_PER_SHAPE} sets of activation buffers; raise it to keep more forwards alive)")
    if plan.done_generation == generation:
        raise RuntimeError("camvid_b200: second backward() through the same forward pass (retain_graph) is not "
                           "supported: the first one reused the saved conv outputs for their gradients")


@torch.library.custom_op("camvid_b200::net_backward", mutates_args=())
def net_backward_op(dlogits: torch.Tensor, plan_handle: int, generation: int) -> torch.Tensor:
    """Gradients of the loss w.r.t. every entry of `params` of the matching net_forward call, as ONE flat fp32 buffer
    (laid out in backward completion order, the unit of the data-parallel all-reduce buckets); Plan.grads_for slices it
    into per-parameter views."""
    plan = _plan_of(plan_handle)
    _check_generation(plan, generation)
    with ops.on_device(dlogits):  # autograd's worker thread of this device, but do not rely on it
        flat = plan.backward(dlogits.float().contiguous())
    plan.done_generation, plan.busy_generation = generation, None
    return flat


@net_backward_op.register_fake
def _net_backward_fake(dlogits, plan_handle, generation):
    return dlogits.new_empty(_plan_of(plan_handle).flat_size)


class _Pending:
    """Lives in the autograd context of a recorded forward: while it is alive a backward may still come, so the plan's
    activation buffers must not be handed to another forward (run_module picks or builds another plan instead). Freed
    with the graph -- `del loss`, the end of the iteration -- it releases the plan."""

    def __init__(self, plan, generation):
        self.plan, self.generation = weakref.ref(plan), generation
        plan.busy_generation = generation

    def __del__(self):
        plan = self.plan()
        if plan is not None and plan.busy_generation == self.generation:
            plan.busy_generation = None


def _net_setup_context(ctx, inputs, output):
    _, _, handle, train = inputs
    ctx.handle, ctx.train = handle, train
    plan = _plan_of(handle)
    ctx.generation = plan.generation
    ctx.pending = _Pending(plan, plan.generation) if train else None


def _net_backward(ctx, dlogits):
    if not ctx.train:
        raise RuntimeError("camvid_b200: backward() through an eval-mode forward is not supported (BatchNorm is folded "
                           "into the conv epilogues and nothing is saved); call module.train() first")
    flat = net_backward_op(dlogits, ctx.handle, ctx.generation)
    return None, _plan_of(ctx.handle).grads_for(flat), None, None


net_forward_op.register_autograd(_net_backward, setup_context=_net_setup_context)


# ---- one reference layer (BasicConv2d / BasicConv / UpSample2d) as a dispatcher op pair: like the network ops, plus the
# gradient w.r.t. the input
@torch.library.custom_op("camvid_b200::block_forward", mutates_args=())
def block_forward_op(x: torch.Tensor, params: List[torch.Tensor], plan_handle: int, train: bool) -> torch.Tensor:
    plan = _plan_of(plan_handle)
    plan.generation += 1
    with ops.on_device(x):
        return plan.forward(x, train)


@block_forward_op.register_fake
def _block_forward_fake(x, params, plan_handle, train):
    p = _plan_of(plan_handle)
    return x.new_empty(x.shape[0], p.class_num, p.ho, p.wo)


@torch.library.custom_op("camvid_b200::block_backward", mutates_args=())
def block_backward_op(dout: torch.Tensor, plan_handle: int, generation: int, need_dx: bool) -> List[torch.Tensor]:
    plan = _plan_of(plan_handle)
    _check_generation(plan, generation)
    with ops.on_device(dout):
        flat, dx = plan.backward(dout.float().contiguous(), need_dx)
    plan.done_generation, plan.busy_generation = generation, None
    return [flat] + ([dx] if need_dx else [])


@block_backward_op.register_fake
def _block_backward_fake(dout, plan_handle, generation, need_dx):
    p = _plan_of(plan_handle)
    out = [dout.new_empty(p.flat_size)]
    if need_dx:
        out.append(dout.new_empty(p.n, p.block.cin, p.h, p.w))
    return out


def _block_setup_context(ctx, inputs, output):
    x, _, handle, train = inputs
    ctx.handle, ctx.train, ctx.need_dx = handle, train, x.requires_grad
    plan = _plan_of(handle)
    ctx.generation = plan.generation
    ctx.pending = _Pending(plan, plan.generation) if train else None


def _block_backward(ctx, dout):
    if not ctx.train:
        raise RuntimeError("camvid_b200: backward() through an eval-mode forward is not supported; call .train() first")
    res = block_backward_op(dout, ctx.handle, ctx.generation, ctx.need_dx)
    return (res[1] if ctx.need_dx else None), _plan_of(ctx.handle).grads_for(res[0]), None, None


block_forward_op.register_autograd(_block_backward, setup_context=_block_setup_context)


# ---- plan cache of a module: per input geometry a small pool of plans (= sets of activation buffers), least recently
# used geometries evicted
PLANS_PER_SHAPE = int(os.environ.get("CVB_PLANS_PER_SHAPE", "2"))  # forwards of one shape that may await their backward
MAX_SHAPES = int(os.environ.get("CVB_MAX_SHAPES", "4"))            # input geometries kept per module


def touch_running_stats(module):
    """Marks every BatchNorm running statistic of a module as modified (a CUDA-graph replay updates them through raw
    pointers without torch noticing)."""
    stats = [b for m in module.modules() if isinstance(m, torch.nn.BatchNorm2d)
             for b in (m.running_mean, m.running_var) if b is not None]
    if stats:
        torch.autograd.graph.increment_version(stats)


def plans_of(module):
    """Every live plan of a drop-in module, most recently used geometry last (introspection: tests, bench.py)."""
    return [p for pool in module.__dict__.get("_plans", {}).values() for p in pool]


def _acquire_plan(module, key, build):
    """The plan a forward of geometry `key` runs on. A plan whose last recorded forward still awaits its backward is
    skipped when another one is free or may be built (up to PLANS_PER_SHAPE per geometry), so two forwards before a
    backward -- a validation pass inside the training step, gradient accumulation over differently shaped inputs, a
    GAN-style double forward -- work; beyond that the oldest pending forward is overwritten and ITS backward raises.
    Geometries not used recently are dropped (MAX_SHAPES; a ragged last batch no longer pins a second 14 GB plan set
    forever) unless a backward is still pending on them."""
    cache = module.__dict__.setdefault("_plans", collections.OrderedDict())
    pool = cache.get(key)
    if pool is not None and any(b.conv.weight.device.index != key[-1] for p in pool for b in p.blocks[:1]):
        pool = None  # the module moved to another device
    if pool is None:
        pool = cache[key] = []
    cache.move_to_end(key)
    plan = next((p for p in pool if p.busy_generation is None), None)
    if plan is None:
        if len(pool) < PLANS_PER_SHAPE:
            # buffer allocation is not part of the traced computation (writer.add_graph traces the first forward)
            tracing = torch._C._get_tracing_state()
            torch._C._set_tracing_state(None)
            try:
                plan = build()
            finally:
                torch._C._set_tracing_state(tracing)
            pool.append(plan)
        else:
            plan = min(pool, key=lambda p: p.generation_time)
    plan.generation_time = next(_CLOCK)
    while len(cache) > MAX_SHAPES:
        victim = next((k for k, v in cache.items() if k != key and all(p.busy_generation is None for p in v)), None)
        if victim is None:
            break
        del cache[victim]
    return plan


_CLOCK = itertools.count(1)


def _check_input(x, channels, what):
    if not x.is_cuda:
        raise RuntimeError("camvid_b200 runs on CUDA (sm_100a) only: move the module and its input to the GPU; "
                           "there is no CPU path")
    if x.dim() != 4 or x.shape[1] != channels:
        raise RuntimeError(f"{what}: expected input [N,{channels},H,W], got {tuple(x.shape)}")


def run_module(module, plan_cls, x):
    """Shared forward of the drop-in modules: x fp32 NCHW CUDA -> logits fp32 NCHW."""
    _check_input(x, module.input_channels, type(module).__name__)
    if x.shape[2] < 32 or x.shape[3] < 32:
        raise RuntimeError("input must be at least 32x32 (five 2x2 poolings)")
    x = x.detach().float().contiguous()
    n, h, w = int(x.shape[0]), int(x.shape[2]), int(x.shape[3])  # concrete even under torch.jit.trace

    def build():
        with ops.on_device(x):
            return plan_cls(module, n, h, w, x.device)

    plan = _acquire_plan(module, (n, h, w, x.device.index), build)
    return net_forward_op(x, plan.param_list(), plan.handle, bool(module.training))


def run_block(layer, conv, bn, x, upsample=False):
    """Forward of ONE reference layer used on its own (BasicConv2d / BasicConv / UpSample2d .forward): fp32 NCHW in,
    fp32 NCHW out, differentiable w.r.t. the input and the layer's parameters."""
    _check_input(x, conv.in_channels, type(layer).__name__)
    if x.shape[2] < 2 or x.shape[3] < 2:
        raise RuntimeError("input must be at least 2x2")
    xin = x.float().contiguous()
    n, h, w = int(x.shape[0]), int(x.shape[2]), int(x.shape[3])

    def build():
        with ops.on_device(x):
            return BlockPlan(layer, conv, bn, upsample, n, h, w, x.device)

    plan = _acquire_plan(layer, (n, h, w, x.device.index), build)
    return block_forward_op(xin, plan.param_list(), plan.handle, bool(layer.training))
