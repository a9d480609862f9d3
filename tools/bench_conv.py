"""Micro-benchmark of one conv configuration through the C ABI.
    python tools/bench_conv.py n h w cin cout [stats=1] [reps=20]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import ops

n, h, w, cin, cout = map(int, sys.argv[1:6])
stats = int(sys.argv[6]) if len(sys.argv) > 6 else 1
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
dev = torch.device("cuda")
x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
wt = torch.randn(cout, cin, 3, 3, device=dev) * (cin * 9) ** -0.5
wp = ops.pack_weights_fprop(wt, 9, cout, cin)
y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
parts = torch.empty(ops.stat_rows(), 2, cout, device=dev) if stats else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.conv3x3(x, wp, y, stat_partials=parts)
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv3x3(x, wp, y, stat_partials=parts)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
med = ts[len(ts) // 2]
fl = 2.0 * 9 * cin * cout * n * h * w
if int(os.environ.get("CVB_DEBUG", "0")) & 2 and parts is not None:
    tr = parts.view(-1)[:10].view(torch.int64).tolist()
    print("trace: total clk", tr[0], "wait tempty", tr[1], "wait fullA", tr[2], "wait fullB", tr[3], "tiles", tr[4])
print(f"conv {n}x{h}x{w} {cin}->{cout} stats={stats} CVB_DEBUG={os.environ.get('CVB_DEBUG', '0')}: median {med * 1e3:.1f} us "
      f"min {ts[0] * 1e3:.1f} us  {fl / med / 1e9:.1f} TFLOP/s")
