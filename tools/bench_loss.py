"""Micro-benchmark of the loss / metric kernels at the bench geometry (16 x 12 x 360 x 480)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import ops

n, c, h, w = 16, 12, 360, 480
dev = torch.device("cuda")
logits = torch.relu(torch.randn(n, c, h, w, device=dev))
target = torch.randint(0, c, (n, h, w), device=dev)
dl = torch.empty_like(logits)
cm = torch.zeros(c, c, dtype=torch.int64, device=dev)
pred = torch.empty(n, h, w, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
px = n * h * w


def timeit(name, fn, nbytes):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    med = sorted(ts)[3]
    print(f"{name:34s} {med * 1e3:8.1f} us  {nbytes / med / 1e6:8.1f} GB/s  ({nbytes / 1e6:.0f} MB)")


timeit("softmax_ce fwd+bwd (fp32 NCHW)", lambda: ops.softmax_ce_nchw(logits, target, -100, True, dl), px * (48 + 8 + 8 + 48))
timeit("softmax_ce fwd only", lambda: ops.softmax_ce_nchw(logits, target, -100, True, None), px * (48 + 8 + 8))
t8 = target.to(torch.uint8)
timeit("softmax_ce fwd+bwd, uint8 labels", lambda: ops.softmax_ce_nchw(logits, t8, -100, True, dl), px * (48 + 1 + 1 + 48))
timeit("argmax + confusion (fp32 NCHW)", lambda: ops.argmax_confusion_nchw(logits, target, cm), px * (48 + 8))
timeit("argmax + confusion + pred", lambda: ops.argmax_confusion_nchw(logits, target, cm, pred), px * (48 + 8 + 8))
timeit("confusion matrix (int64 labels)", lambda: ops.confusion_matrix(pred, target, c, cm), px * 16)
