"""Per-call kernel times of one training step (CUDA events around every C-ABI call), averaged over a few steps.
    python tools/layer_times.py [unet|segnet] [batch] [h] [w]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import engine, ops

engine.OVERLAP_WGRAD = False  # one stream: unambiguous event brackets
from camvid_b200.nn import CrossEntropyLoss
from camvid_b200.utils import get_model

name = sys.argv[1] if len(sys.argv) > 1 else "unet"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H = int(sys.argv[3]) if len(sys.argv) > 3 else 360
W = int(sys.argv[4]) if len(sys.argv) > 4 else 480
torch.manual_seed(0)
net = get_model(name, 3, 12).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=5e-4)
x = torch.randn(B, 3, H, W, device="cuda")
t = torch.randint(0, 12, (B, H, W), device="cuda")
loss_fn = CrossEntropyLoss()


def step():
    opt.zero_grad(set_to_none=True)
    loss_fn(net(x), t).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
S = 3
rec = ops.profile(True)
for _ in range(S):
    step()
torch.cuda.synchronize()
ops.profile(False)
per = len(rec) // S
tot = 0.0
rows = []
for i in range(per):
    ms = sum(rec[i + k * per][2].elapsed_time(rec[i + k * per][3]) for k in range(S)) / S
    what, work = rec[i][0], rec[i][1]
    tot += ms
    rate = work[1] / (ms * 1e-3) / (1e12 if work[0] == "flops" else 1e9)
    rows.append((i, what, ms * 1e3, rate, "TF/s" if work[0] == "flops" else "GB/s", work[2] if len(work) > 2 else ""))
for r in rows:
    print(f"{r[0]:4d} {r[1]:24s} {r[2]:9.1f} us {r[3]:9.1f} {r[4]} {r[5]}")
print(f"sum of kernel times {tot:.2f} ms/step")
