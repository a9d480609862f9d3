"""Timeline of one training step with the weight-gradient side stream ON: start / end of every C-ABI call relative to the
step's first call (CUDA events on the stream each call was enqueued on). Shows what actually overlaps.
    python tools/timeline.py [unet|segnet] [batch] [first] [last]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import ops
from camvid_b200.nn import CrossEntropyLoss
from camvid_b200.utils import get_model

name = sys.argv[1] if len(sys.argv) > 1 else "unet"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
last = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
torch.manual_seed(0)
net = get_model(name, 3, 12).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=5e-4)
x = torch.randn(B, 3, 360, 480, device="cuda")
t = torch.randint(0, 12, (B, 360, 480), device="cuda")
loss_fn = CrossEntropyLoss()


def step():
    opt.zero_grad(set_to_none=True)
    loss_fn(net(x), t).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
rec = ops.profile(True)
step()
torch.cuda.synchronize()
ops.profile(False)
base = rec[0][2]
for i, (what, work, e0, e1) in enumerate(rec):
    if first <= i <= last:
        shape = work[2] if len(work) > 2 else ""
        print(f"{i:4d} {what:24s} {base.elapsed_time(e0) * 1e3:9.1f} -> {base.elapsed_time(e1) * 1e3:9.1f} us  "
              f"({e0.elapsed_time(e1) * 1e3:7.1f}) {shape}")
