"""Drop-in for the loss the reference trains with: nn.CrossEntropyLoss() on fp32 NCHW logits and int64 targets
(train.py:105,130-131; eval.py:42,58), backed by the fused softmax / loss / gradient kernel."""
import torch

from . import ops


class _CEFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, logits, target, ignore_index, reduction):
        if not (logits.is_cuda and logits.dim() == 4):
            raise RuntimeError("camvid_b200.nn.CrossEntropyLoss expects CUDA logits of shape [N,C,H,W]")
        if target.dtype != torch.int64 or tuple(target.shape) != (logits.shape[0], logits.shape[2], logits.shape[3]):
            raise RuntimeError("expected an int64 target of shape [N,H,W]")
        lg = logits.detach().float().contiguous()
        tg = target.contiguous()
        acc = torch.zeros(2, dtype=torch.float64, device=lg.device)
        need_grad = logits.requires_grad
        dl = torch.empty_like(lg) if need_grad else None
        if reduction == "mean":
            if 0 <= ignore_index < lg.shape[1]:
                inv = (1.0 / (tg != ignore_index).sum().to(torch.float32)).reshape(1)
                ops.softmax_ce_nchw(lg, tg, ignore_index, acc, dl, 1.0, inv)
            else:
                ops.softmax_ce_nchw(lg, tg, ignore_index, acc, dl, 1.0 / tg.numel())
            loss = (acc[0] / acc[1]).to(torch.float32)
        else:
            ops.softmax_ce_nchw(lg, tg, ignore_index, acc, dl, 1.0)
            loss = acc[0].to(torch.float32)
        ctx.dl = dl
        return loss

    @staticmethod
    def backward(ctx, g):
        dl = ctx.dl
        ctx.dl = None
        return dl.mul_(g), None, None, None


class CrossEntropyLoss(torch.nn.Module):
    """Supports what the reference uses: no class weights, no label smoothing, reduction 'mean' (default) or 'sum'."""

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
