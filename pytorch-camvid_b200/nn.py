"""Drop-in for the loss the reference trains with: nn.CrossEntropyLoss() on fp32 NCHW logits and int64 targets
(train.py:105,130-131; eval.py:42,58), backed by the fused softmax / loss / gradient kernel. Stock
torch.nn.CrossEntropyLoss works on the drop-in modules' logits too (they are ordinary autograd tensors); this one
saves the separate log_softmax / nll_loss passes and also accepts the uint8 masks camvid_b200.data keeps on the device.
"""
import torch

from . import ops


class _CEFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, target, ignore_index, reduction):
        if not (logits.is_cuda and logits.dim() == 4):
            raise RuntimeError("camvid_b200.nn.CrossEntropyLoss expects CUDA logits of shape [N,C,H,W]")
        if target.dtype not in (torch.int64, torch.uint8) or \
                tuple(target.shape) != (logits.shape[0], logits.shape[2], logits.shape[3]):
            raise RuntimeError("expected an int64 (or uint8) target of shape [N,H,W]")
        if target.device != logits.device:
            raise RuntimeError("logits and target are on different devices")
        with ops.on_device(logits):
            lg = logits.detach().float().contiguous()
            tg = target.contiguous()
            dl = torch.empty_like(lg) if logits.requires_grad else None
            # the counted-pixel total behind reduction='mean' is taken on the device inside the call (pre-pass over
            # the target), so the gradient scale always matches the loss denominator -- ignore_index in or out of range
            loss = ops.softmax_ce_nchw(lg, tg, ignore_index, reduction == "mean", dl)
        ctx.dl = dl
        return loss

    @staticmethod
    def backward(ctx, g):
        dl = ctx.dl
        ctx.dl = None  # the gradient was computed with the loss; it is handed over (scaled in place), not kept
        if dl is None:
            raise RuntimeError("camvid_b200.nn.CrossEntropyLoss: second backward() through the same loss value is not "
                               "supported (its gradient buffer was released by the first)")
        return dl.mul_(g), None, None, None


class CrossEntropyLoss(torch.nn.Module):
    """Supports what the reference uses: no class weights, no label smoothing, reduction 'mean' (default) or 'sum'.
    A target that is neither ignore_index nor in [0, C) makes the loss NaN (torch raises a device-side assert)."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction='mean',
                 label_smoothing=0.0):
        super().__init__()
        if weight is not None or label_smoothing != 0.0 or reduction not in ('mean', 'sum') \
                or size_average is not None or reduce is not None:
            raise ValueError("camvid_b200.nn.CrossEntropyLoss supports weight=None, label_smoothing=0 and "
                             "reduction in ('mean', 'sum') only")
        self.ignore_index, self.reduction = ignore_index, reduction

    def forward(self, input, target):
        return _CEFunction.apply(input, target, self.ignore_index, self.reduction)
