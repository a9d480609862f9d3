"""Host-side mirror of the reference's legacy/metrics.py::Metrics (legacy/metrics.py:6-71).

Same constructor, methods and numbers; `add` counts on the GPU (C ABI cvb_confusion_matrix, rows = ground truth,
columns = prediction, pairs with a label outside range(class_num) dropped like sklearn's `labels=range(C)`) and also
accepts CUDA tensors, so eval.py:62-64 no longer needs the device->host copy of every prediction map.
"""
import numpy as np
import torch

from .. import ops
from ..utils import _to_cuda_i64


class Metrics:

    def __init__(self, class_num, ignore_index=None):
        self.class_num = class_num
        self.ignore_index = ignore_index
        self._confusion_matrix = np.zeros((self.class_num, self.class_num))

    def add(self, preds, gts):
        """preds, gts: 1-D label arrays / tensors of equal length (legacy/metrics.py:22-30)."""
        p, g = _to_cuda_i64(preds).view(-1), _to_cuda_i64(gts).view(-1)
        if p.numel() != g.numel():
            raise ValueError("Found input variables with inconsistent numbers of samples: [%d, %d]"
                             % (g.numel(), p.numel()))
        if g.device != p.device:
            g = g.to(p.device)
        with ops.on_device(p):
            cm = torch.zeros(self.class_num, self.class_num, dtype=torch.int64, device=p.device)
            ops.confusion_matrix(p, g, self.class_num, cm)
        self._confusion_matrix += cm.cpu().numpy()

    def add_logits(self, logits, gts):
        """Extension: fused argmax(dim=1) + counting straight from fp32 NCHW logits (eval.py:61-64 in one kernel)."""
        if not (torch.is_tensor(gts) and gts.is_cuda and gts.dtype == torch.uint8):  # uint8 device masks pass through
            gts = _to_cuda_i64(gts)
        g = gts.to(logits.device).contiguous()
        with ops.on_device(logits):
            cm = torch.zeros(self.class_num, self.class_num, dtype=torch.int64, device=g.device)
            ops.argmax_confusion_nchw(logits.detach().float().contiguous(), g, cm)
        self._confusion_matrix += cm.cpu().numpy()

    def clear(self):
        self._confusion_matrix.fill(0)

    def _keep(self):
        return [i for i in range(self.class_num) if i != self.ignore_index]

    def _ratio(self, denom, drop_ignored, average):
        tp = np.diag(self._confusion_matrix)
        val = tp / (denom + 1e-15)
        if drop_ignored:
            val = val[self._keep()]
        return val.mean() if average else val

    def precision(self, average=True):
        # legacy/metrics.py:35-46: the ignore class is dropped only when `ignore_index` is truthy (so not for 0/None)
        return self._ratio(self._confusion_matrix.sum(axis=0), bool(self.ignore_index), average)

    def recall(self, average=True):
        # legacy/metrics.py:48-59
        return self._ratio(self._confusion_matrix.sum(axis=1), bool(self.ignore_index), average)

    def iou(self, average=True):
        # legacy/metrics.py:61-71: always drops index == ignore_index (a None never matches)
        cm = self._confusion_matrix
        return self._ratio(cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm), True, average)
