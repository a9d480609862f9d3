"""Thin operator layer over the C ABI: torch tensors in, kernels enqueued on torch's current CUDA stream.

Activations are torch bf16 tensors of logical shape [N, H, W, C] with stride(3) == 1 (NHWC); channel slices and row
windows of a larger buffer are passed as strided views, no copies. Nothing here computes on the host or falls back to
PyTorch kernels: every function ends in one call into libcamvid_b200.so and raises RuntimeError on failure.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvEpilogue, View


LAUNCHES = 0  # kernels enqueued through the C ABI by this process (bench.py reports the per-step count)
_PROFILE = None  # while kernel timing is on: list of (entry point, algorithmic work, start event, end event)
WORK_SCALE = 1.0  # real / padded channel ratio of the block being run (engine sets it; algorithmic-byte accounting)


def profile(on=True):
    """Turns per-call CUDA-event timing on (returns the record list) or off. Measurement aid of bench.py: events are
    recorded on torch's current stream, which is the stream every kernel is enqueued on."""
    global _PROFILE
    _PROFILE = [] if on else None
    return _PROFILE


def _call(what, kernels, work, fn, *args):
    """One C-ABI call: counts its kernel launches, optionally brackets it with CUDA events, raises on failure.
    work = ("bytes" | "flops", algorithmic amount) for the roofline arithmetic."""
    global LAUNCHES
    LAUNCHES += kernels
    prof = _PROFILE
    if prof is None:
        rc = fn(*args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        prof.append((what, work, e0, e1))
    if rc != 0:
        _lib.check(rc, what)


def _nbytes(*ts):
    """Algorithmic bytes of dense passes over these tensors (bf16 views count their real channels)."""
    tot = 0.0
    for t in ts:
        if t is not None:
            tot += t.numel() * t.element_size() * (WORK_SCALE if t.dtype == torch.bfloat16 else 1.0)
    return ("bytes", tot)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(t):
    """Context manager making t's GPU the current device: kernels are enqueued on the CURRENT device's current stream
    and events / SM counts are taken from it, so every public entry point (module forward / backward, loss, metrics,
    optimizer) wraps its work in this -- a module on cuda:1 works while cuda:0 is current."""
    if not t.is_cuda:
        raise RuntimeError("camvid_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    return torch.cuda.device(t.device)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def view(t):
    """cvb_view of an NHWC bf16 tensor (possibly a strided slice)."""
    if t is None:
        return View(None, 0, 0, 0, 0, 0, 0, 0)
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4):
        raise RuntimeError(f"expected a CUDA bf16 [N,H,W,C] tensor, got {t.dtype} {tuple(t.shape)} on {t.device}")
    if t.stride(3) != 1:
        raise RuntimeError("NHWC view must have channel stride 1")
    n, h, w, c = t.shape
    return View(t.data_ptr(), n, h, w, c, t.stride(0), t.stride(1), t.stride(2))


def _f32(t, name):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"{name}: expected a contiguous CUDA fp32 tensor")


def pad64(c):
    return (c + 63) // 64 * 64


def sm_count():
    return _lib.load().cvb_sm_count()


def stat_rows():
    return _lib.load().cvb_conv_stat_rows()


# ---------------------------------------------------------------- layout
def nchw_to_nhwc(src, dst):
    """src fp32 [N,C,H,W] contiguous -> dst NHWC bf16 view (extra channels zeroed)."""
    _f32(src, "nchw_to_nhwc.src")
    _call("nchw_to_nhwc", 1, _nbytes(src, dst), _lib.load().cvb_nchw_f32_to_nhwc_bf16, _ptr(src), src.shape[1],
          view(dst), _stream())
    return dst


def nhwc_to_nchw(src, dst):
    """src NHWC bf16 view -> dst fp32 [N,C,H,W] contiguous (first C channels)."""
    _f32(dst, "nhwc_to_nchw.dst")
    _call("nhwc_to_nchw", 1, _nbytes(src, dst), _lib.load().cvb_nhwc_bf16_to_nchw_f32, view(src), _ptr(dst),
          dst.shape[1], _stream())
    return dst


def im2col3x3(src, dst):
    _f32(src, "im2col3x3.src")
    _call("im2col3x3", 1, _nbytes(src, dst), _lib.load().cvb_im2col3x3_nchw_f32, _ptr(src), src.shape[1], view(dst),
          _stream())
    return dst


def zero_view(t):
    _call("zero_view", 1, _nbytes(t), _lib.load().cvb_zero_view, view(t), _stream())
    return t


def input_stage_u8(img_u8, mean, std, out, mask_u8=None, mask_i64=None):
    """transforms.ToTensor + transforms.Normalize (transforms.py:485-538) on the device: img_u8 uint8 [N,H,W,C] (HWC, as
    cv2 yields it) -> out fp32 [N,C,H,W], bit-exact with the reference transforms; mask_u8 uint8 [N,H,W] -> mask_i64
    (optional). Either half may be None."""
    if img_u8 is not None:
        if not (img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.is_contiguous()):
            raise RuntimeError("input_stage_u8: expected a contiguous CUDA uint8 [N,H,W,C] image batch")
        n, h, w, c = img_u8.shape
        _f32(out, "input_stage_u8.out")
        if tuple(out.shape) != (n, c, h, w) or len(mean) != c or len(std) != c:
            raise RuntimeError("input_stage_u8: out must be [N,C,H,W] and mean / std must have C entries")
    else:
        n, h, w = mask_u8.shape
        c = 0
    if mask_i64 is not None:
        if not (mask_u8 is not None and mask_u8.is_cuda and mask_u8.dtype == torch.uint8 and mask_u8.is_contiguous()
                and mask_i64.is_cuda and mask_i64.dtype == torch.int64 and mask_i64.is_contiguous()
                and tuple(mask_u8.shape) == (n, h, w) == tuple(mask_i64.shape)):
            raise RuntimeError("input_stage_u8: masks must be contiguous CUDA uint8 / int64 [N,H,W] tensors")
    fl = ctypes.c_float * max(c, 1)
    _call("input_stage_u8", 1, _nbytes(img_u8, out, mask_u8 if mask_i64 is not None else None, mask_i64),
          _lib.load().cvb_input_stage_u8, _ptr(img_u8), n, h, w, c, fl(*[float(v) for v in mean][:c]) if c else None,
          fl(*[float(v) for v in std][:c]) if c else None, _ptr(out), _ptr(mask_u8 if mask_i64 is not None else None),
          _ptr(mask_i64), _stream())
    return out, mask_i64


# ---------------------------------------------------------------- convolution
def pack_weights_fprop(w, taps, cout_pad, cin_pad, out=None):
    """OIHW fp32 -> bf16 [cout_pad, taps*cin_pad] GEMM-B matrix."""
    _f32(w, "pack_weights_fprop.w")
    cout, cin = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty(cout_pad, taps * cin_pad, dtype=torch.bfloat16, device=w.device)
    _call("pack_weights_fprop", 1, _nbytes(w, out), _lib.load().cvb_pack_weights_fprop, _ptr(w), cout, cin, taps,
          cout_pad, cin_pad, _ptr(out), _stream())
    return out


def pack_weights_dgrad(w, cout_pad, cin_pad, out=None):
    """OIHW fp32 -> bf16 [cin_pad, 9*cout_pad] rotated/transposed GEMM-B matrix of the data-gradient conv."""
    _f32(w, "pack_weights_dgrad.w")
    cout, cin = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty(cin_pad, 9 * cout_pad, dtype=torch.bfloat16, device=w.device)
    _call("pack_weights_dgrad", 1, _nbytes(w, out), _lib.load().cvb_pack_weights_dgrad, _ptr(w), cout, cin, cout_pad,
          cin_pad, _ptr(out), _stream())
    return out


def pack_table(entries, device):
    """Device table for pack_weights_batch. entries: (w fp32 OIHW, dst_fprop, dst_dgrad or None, cout_pad, cin_pad)."""
    rows = []
    for w, df, dd, cout_pad, cin_pad in entries:
        _f32(w, "pack_table.w")
        rows.append([w.data_ptr(), df.data_ptr(), dd.data_ptr() if dd is not None else 0, w.shape[0], w.shape[1],
                     cout_pad, cin_pad, 0])
    return torch.tensor(rows, dtype=torch.int64).to(device)


def pack_weights_batch(table, max_cout_pad, max_cin_pad, nbytes):
    """One launch packing every 3x3 layer's fprop + dgrad operands (cvb_pack_entry table on the device)."""
    _call("pack_weights_batch", 1, ("bytes", float(nbytes)), _lib.load().cvb_pack_weights_batch, _ptr(table),
          table.shape[0], max_cout_pad, max_cin_pad, _stream())


def conv3x3(x, wpack, y, taps=9, stat_partials=None, scale=None, shift=None, relu=False, algo_flops=None, bwd=None):
    """y = conv(x, wpack). Optional epilogues: BN statistics partials (train), folded scale/shift(+ReLU) (eval), or --
    for a data gradient, bwd = (y_prev, scale_prev, shift_prev, partials) -- the BatchNorm+ReLU backward reduction of
    the block that consumes this gradient (only where conv3x3_fuses_bwd_stats(x, y) is True).
    algo_flops: FLOPs of the un-padded convolution (roofline accounting); default = the padded GEMM's."""
    ep = ConvEpilogue(stat_partials.data_ptr() if stat_partials is not None else None,
                      scale.data_ptr() if scale is not None else None,
                      shift.data_ptr() if shift is not None else None, 1 if relu else 0,
                      view(None), None, None, None)
    if bwd is not None:
        by, bscale, bshift, bparts = bwd
        _f32(bparts, "conv3x3.bwd_partials")
        if bparts.numel() < stat_rows() * 2 * y.shape[3]:
            raise RuntimeError("conv3x3: bwd partials too small")
        ep.bwd_y, ep.bwd_scale, ep.bwd_shift, ep.bwd_partials = view(by), bscale.data_ptr(), bshift.data_ptr(), bparts.data_ptr()
    if stat_partials is not None:
        _f32(stat_partials, "conv3x3.stat_partials")
        if stat_partials.numel() < stat_rows() * 2 * y.shape[3]:
            raise RuntimeError("conv3x3: stat_partials too small")
    if algo_flops is None:
        algo_flops = 2.0 * taps * x.shape[3] * y.shape[3] * y.shape[0] * y.shape[1] * y.shape[2]
    _call("conv3x3_fprop", 1, ("flops", algo_flops, f"{tuple(x.shape)}->{y.shape[3]} taps{taps}",
                               2.0 * (x.numel() + y.numel() + wpack.numel())), _lib.load().cvb_conv3x3_fprop, view(x), _ptr(wpack), taps,
          view(y), ctypes.byref(ep), _stream())
    return y


def conv3x3_fuses_bwd_stats(x, y, taps=9):
    return bool(_lib.load().cvb_conv3x3_fprop_fuses_bwd_stats(view(x), view(y), taps))


def conv3x3_wgrad_workspace_bytes(x, dy, taps=9):
    r = _lib.load().cvb_conv3x3_wgrad_workspace_bytes(view(x), view(dy), taps)
    if r < 0:
        _lib.check(int(r), "conv3x3_wgrad_workspace_bytes")
    return int(r)


def conv3x3_wgrad(x, dy, dw, taps=9, workspace=None, algo_flops=None):
    """dw (fp32 OIHW [cout,cin,3,3], overwritten) = weight gradient from activations x and output gradient dy."""
    _f32(dw, "conv3x3_wgrad.dw")
    cout, cin = dw.shape[0], dw.shape[1]
    if workspace is None:
        workspace = torch.empty(conv3x3_wgrad_workspace_bytes(x, dy, taps), dtype=torch.uint8, device=x.device)
    if algo_flops is None:
        algo_flops = 2.0 * taps * x.shape[3] * dy.shape[3] * dy.shape[0] * dy.shape[1] * dy.shape[2]
    _call("conv3x3_wgrad", 2, ("flops", algo_flops, f"{tuple(x.shape)}->{dy.shape[3]} taps{taps}"), _lib.load().cvb_conv3x3_wgrad, view(x), view(dy), taps, _ptr(dw),
          cout, cin, _ptr(workspace), workspace.numel() * workspace.element_size(), _stream())
    return dw


# ---------------------------------------------------------------- batch norm (+ReLU)
def bn_stats(y, partials, rows):
    _call("bn_stats", 1, _nbytes(y), _lib.load().cvb_bn_stats, view(y), _ptr(partials), rows, _stream())
    return partials


def bn_finalize(partials, rows, c, c_pad, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps,
                mean, invstd, scale, shift):
    _call("bn_finalize", 1, ("bytes", rows * 2 * c_pad * 4.0), _lib.load().cvb_bn_finalize, _ptr(partials), rows, c,
          c_pad, count, _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(running_mean), _ptr(running_var), momentum,
          eps, _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), _stream())


def bn_relu_apply(y, scale, shift, a, reverse=False):
    """reverse: walk the tensor back to front (same result; see cvb_bn_relu_apply in the header for when that pays)."""
    _call("bn_relu_apply", 1, _nbytes(y, a), _lib.load().cvb_bn_relu_apply, view(y), _ptr(scale), _ptr(shift),
          view(a), int(reverse), _stream())
    return a


def bn_relu_bwd_reduce(da, y, scale, shift, partials, rows, reverse=False):
    _call("bn_relu_bwd_reduce", 1, _nbytes(da, y), _lib.load().cvb_bn_relu_bwd_reduce, view(da), view(y), _ptr(scale),
          _ptr(shift), _ptr(partials), rows, int(reverse), _stream())


def bn_relu_apply_nchw(y, scale, shift, dst):
    """Last block + module boundary: dst fp32 [N,C,H,W] = float(bf16(relu(y*scale+shift))) from the NHWC bf16 view y."""
    _f32(dst, "bn_relu_apply_nchw.dst")
    _call("bn_relu_apply_nchw", 1, _nbytes(y, dst), _lib.load().cvb_bn_relu_apply_nchw_f32, view(y), _ptr(scale),
          _ptr(shift), _ptr(dst), dst.shape[1], _stream())
    return dst


def nchw_to_nhwc_bn_reduce(src, da, y, scale, shift, partials, rows):
    """Entry of the last block's backward: da = bf16 NHWC of the fp32 NCHW gradient `src` + the BatchNorm+ReLU backward
    reduction over (da, y) in the same pass."""
    _f32(src, "nchw_to_nhwc_bn_reduce.src")
    _call("nchw_to_nhwc_bn_reduce", 1, _nbytes(src, da, y), _lib.load().cvb_nchw_f32_to_nhwc_bf16_bn_reduce, _ptr(src),
          src.shape[1], view(da), view(y), _ptr(scale), _ptr(shift), _ptr(partials), rows, _stream())


def bn_bwd_finalize(partials, rows, c, c_pad, count, gamma, mean, invstd, dgamma, dbeta, coef):
    _call("bn_bwd_finalize", 1, ("bytes", rows * 2 * c_pad * 4.0), _lib.load().cvb_bn_bwd_finalize, _ptr(partials),
          rows, c, c_pad, count, _ptr(gamma), _ptr(mean), _ptr(invstd), _ptr(dgamma), _ptr(dbeta), _ptr(coef),
          _stream())


def bn_relu_bwd_apply(da, y, scale, shift, coef, dy, reverse=False):
    _call("bn_relu_bwd_apply", 1, _nbytes(da, y, dy), _lib.load().cvb_bn_relu_bwd_apply, view(da), view(y),
          _ptr(scale), _ptr(shift), _ptr(coef), view(dy), int(reverse), _stream())
    return dy


# ---------------------------------------------------------------- pooling
def maxpool2x2(x, out, code=None):
    _call("maxpool2x2_fwd", 1, _nbytes(x, out, code), _lib.load().cvb_maxpool2x2_fwd, view(x), view(out), _ptr(code),
          _stream())
    return out


def bn_relu_maxpool2x2(y, scale, shift, a, out, code=None):
    _call("bn_relu_maxpool2x2_fwd", 1, _nbytes(y, a, out, code), _lib.load().cvb_bn_relu_maxpool2x2_fwd, view(y),
          _ptr(scale), _ptr(shift), view(a), view(out), _ptr(code), _stream())
    return out


def maxpool2x2_bwd(dout, dx, code=None, x=None, accumulate=False):
    _call("maxpool2x2_bwd", 1, _nbytes(dout, dx, code, x, dx if accumulate else None),
          _lib.load().cvb_maxpool2x2_bwd, view(dout), _ptr(code), view(x), view(dx), 1 if accumulate else 0,
          _stream())
    return dx


def maxpool2x2_bwd_bn_reduce(dout, dx, y, scale, shift, partials, rows, code=None, accumulate=False):
    """maxpool2x2_bwd + bn_relu_bwd_reduce of the pooled block in one pass over dx (see the header)."""
    _call("maxpool2x2_bwd_bn_reduce", 1, _nbytes(dout, dx, code, y, dx if accumulate else None),
          _lib.load().cvb_maxpool2x2_bwd_bn_reduce, view(dout), _ptr(code), view(y), _ptr(scale), _ptr(shift), view(dx),
          1 if accumulate else 0, _ptr(partials), rows, _stream())
    return dx


def maxunpool2x2(x, code, out):
    _call("maxunpool2x2_fwd", 1, _nbytes(x, code, out), _lib.load().cvb_maxunpool2x2_fwd, view(x), _ptr(code),
          view(out), _stream())
    return out


def maxunpool2x2_bwd(dout, code, dx):
    # reads one of the four window positions per output element: a quarter of dout is touched algorithmically
    _call("maxunpool2x2_bwd", 1, _nbytes(dx, dx, code), _lib.load().cvb_maxunpool2x2_bwd, view(dout), _ptr(code),
          view(dx), _stream())
    return dx


def pool_code_to_index(code, w_in):
    """uint8 codes [N,Ho,Wo,C] -> torch-style int64 indices [N,C,Ho,Wo] (h*W_in + w per plane)."""
    n, ho, wo, c = code.shape
    idx = torch.empty(n, c, ho, wo, dtype=torch.int64, device=code.device)
    _call("pool_code_to_index", 1, _nbytes(code, idx), _lib.load().cvb_pool_code_to_index, _ptr(code), n, ho, wo, c,
          w_in, _ptr(idx), _stream())
    return idx


# ---------------------------------------------------------------- upsample
def bilinear2x(x, out):
    _call("bilinear2x_fwd", 1, _nbytes(x, out), _lib.load().cvb_bilinear2x_fwd, view(x), view(out), _stream())
    return out


def bn_relu_bilinear2x(y, scale, shift, out):
    """out = upsample2x(bf16(relu(y*scale + shift))): the producer block's BatchNorm+ReLU fused into the upsampling."""
    _call("bn_relu_bilinear2x_fwd", 1, _nbytes(y, out), _lib.load().cvb_bn_relu_bilinear2x_fwd, view(y), _ptr(scale),
          _ptr(shift), view(out), _stream())
    return out


def bilinear2x_bwd(dout, dx):
    _call("bilinear2x_bwd", 1, _nbytes(dout, dx), _lib.load().cvb_bilinear2x_bwd, view(dout), view(dx), _stream())
    return dx


# ---------------------------------------------------------------- loss / metric
def label_type(t, name):
    """cvb_label_type of a label tensor: int64 (the reference API) or uint8 (device-resident masks of the input stage)."""
    if not (t.is_cuda and t.is_contiguous() and t.dtype in (torch.int64, torch.uint8)):
        raise RuntimeError(f"{name}: expected a contiguous CUDA int64 or uint8 label tensor, got {t.dtype} on {t.device}")
    return 8 if t.dtype == torch.int64 else 1


def softmax_ce_nchw(logits, target, ignore_index, mean, dlogits, grad_scale=1.0):
    """Fused CrossEntropyLoss forward (+ gradient when dlogits is given). Returns the loss as a device scalar (fp32);
    NaN if a label is neither ignore_index nor in [0, C). Two launches for reduction='mean' (counted-pixel pre-pass)."""
    _f32(logits, "softmax_ce.logits")
    n, c, h, w = logits.shape
    scratch = torch.zeros(4, dtype=torch.float64, device=logits.device)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    _call("softmax_ce_nchw_f32", 2 if mean else 1, _nbytes(logits, target, dlogits), _lib.load().cvb_softmax_ce_nchw_f32,
          _ptr(logits), _ptr(target), label_type(target, "softmax_ce.target"), n, c, h, w, ignore_index,
          1 if mean else 0, _ptr(scratch), _ptr(loss), _ptr(dlogits), grad_scale, _stream())
    return loss


def softmax_ce_nhwc(logits, c, target, ignore_index, mean, dlogits, grad_scale=1.0):
    px = logits.shape[0] * logits.shape[1] * logits.shape[2]
    scratch = torch.zeros(4, dtype=torch.float64, device=logits.device)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    _call("softmax_ce_nhwc_bf16", 2 if mean else 1,
          ("bytes", px * (c * 2 * (2 if dlogits is not None else 1) + float(target.element_size()))),
          _lib.load().cvb_softmax_ce_nhwc_bf16, view(logits), c, _ptr(target), label_type(target, "softmax_ce.target"),
          ignore_index, 1 if mean else 0, _ptr(scratch), _ptr(loss), view(dlogits), grad_scale, _stream())
    return loss


NO_IGNORE = -(2 ** 62)


def confusion_matrix(pred, gt, c, cm, ignore_label=NO_IGNORE, clamp_oob=False):
    """cm[gt, pred] += counts over label tensors (any shape, same numel, same dtype: int64 or uint8)."""
    lt = label_type(pred, "confusion_matrix.pred")
    if label_type(gt, "confusion_matrix.gt") != lt:
        raise RuntimeError("confusion_matrix: pred and gt must have the same dtype")
    if pred.numel() != gt.numel():
        raise RuntimeError("confusion_matrix: pred and gt sizes differ")
    _call("confusion_matrix", 1, _nbytes(pred, gt), _lib.load().cvb_confusion_matrix, _ptr(pred), _ptr(gt), lt,
          pred.numel(), c, ignore_label, 1 if clamp_oob else 0, _ptr(cm), _stream())
    return cm


def argmax_confusion_nchw(logits, gt, cm, pred=None):
    _f32(logits, "argmax_confusion.logits")
    n, c, h, w = logits.shape
    _call("argmax_confusion_nchw_f32", 1, _nbytes(logits, gt, pred), _lib.load().cvb_argmax_confusion_nchw_f32,
          _ptr(logits), _ptr(gt), label_type(gt, "argmax_confusion.gt"), n, c, h, w, _ptr(pred), _ptr(cm), _stream())
    return cm


def argmax_confusion_nhwc(logits, c, gt, cm, pred=None):
    px = logits.shape[0] * logits.shape[1] * logits.shape[2]
    _call("argmax_confusion_nhwc_bf16", 1, ("bytes", px * (c * 2 + float(gt.element_size()) + (8.0 if pred is not None else 0.0))),
          _lib.load().cvb_argmax_confusion_nhwc_bf16, view(logits), c, _ptr(gt), label_type(gt, "argmax_confusion.gt"),
          _ptr(pred), _ptr(cm), _stream())
    return cm
