"""Micro-benchmark of one weight-gradient configuration through the C ABI (+ check against torch).
    python tools/bench_wgrad.py n h w cin cout [reps=10]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import ops

n, h, w, cin, cout = map(int, sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 10
dev = torch.device("cuda")
x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
dy = torch.randn(n, h, w, cout, device=dev).to(torch.bfloat16)
dw = torch.empty(cout, cin, 3, 3, device=dev)
ws = torch.empty(ops.conv3x3_wgrad_workspace_bytes(x, dy), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    ops.conv3x3_wgrad(x, dy, dw, workspace=ws)
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3_wgrad(x, dy, dw, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
med = ts[len(ts) // 2]
fl = 2.0 * 9 * cin * cout * n * h * w
msg = ""
if n * h * w * cin * cout <= 16 * 90 * 120 * 256 * 256:
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    msg = f" rel err {((dw - ref).norm() / ref.norm()).item():.2e}"
print(f"wgrad {n}x{h}x{w} {cin}->{cout}: median {med * 1e3:.1f} us  {fl / med / 1e9:.1f} TFLOP/s  ws {ws.numel() / 2**20:.1f} MiB{msg}")
