"""Micro-benchmark of the BatchNorm / pool / bilinear memory-bound kernels on one tensor shape.
    python tools/bench_bn.py n h w c"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import ops

n, h, w, c = map(int, sys.argv[1:5])
dev = torch.device("cuda")
y = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
da = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
a = torch.empty_like(y)
dy = torch.empty_like(y)
pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=dev)
up = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=dev) if n * h * w * c < 2 ** 28 else None
scale = torch.rand(c, device=dev) + 0.5
shift = torch.randn(c, device=dev)
coef = torch.randn(3, c, device=dev)
rows = 4 * ops.sm_count()
parts = torch.empty(rows, 2, c, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
E = y.numel()


def timeit(name, fn, nbytes, reps=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name:22s} {med * 1e3:8.1f} us  {nbytes / med / 1e6:8.1f} GB/s  ({nbytes / 1e6:.0f} MB)")


print(f"shape {n}x{h}x{w}x{c} CVB_EW_PER_SM={os.environ.get('CVB_EW_PER_SM', '-')}")
timeit("bn_relu_apply", lambda: ops.bn_relu_apply(y, scale, shift, a), 4 * E)
timeit("bn_relu_bwd_reduce", lambda: ops.bn_relu_bwd_reduce(da, y, scale, shift, parts, rows), 4 * E)
timeit("bn_relu_bwd_apply", lambda: ops.bn_relu_bwd_apply(da, y, scale, shift, coef, dy), 6 * E)
timeit("bn_relu_maxpool", lambda: ops.bn_relu_maxpool2x2(y, scale, shift, a, pooled), 4 * E + E // 2)
timeit("maxpool_bwd(recompute)", lambda: ops.maxpool2x2_bwd(pooled, dy, x=a, accumulate=True), 6 * E + E // 2)
da2 = da.clone()
code8 = torch.randint(0, 4, pooled.shape, dtype=torch.uint8, device=dev)
timeit("pool_bwd+reduce (2 kernels)", lambda: (ops.maxpool2x2_bwd(pooled, da2, x=a, accumulate=True),
                                              ops.bn_relu_bwd_reduce(da2, y, scale, shift, parts, rows)), 10 * E + E // 2)
timeit("pool_bwd_bn_reduce fused", lambda: ops.maxpool2x2_bwd_bn_reduce(pooled, da2, y, scale, shift, parts, rows, code=code8, accumulate=True),
       6 * E + E // 2)
if up is not None:
    timeit("bilinear2x_fwd", lambda: ops.bilinear2x(y, up), 2 * E + 8 * E)
    timeit("bilinear2x_bwd", lambda: ops.bilinear2x_bwd(up, dy), 2 * E + 8 * E)
