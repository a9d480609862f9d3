"""Interleaved in-process comparison of the visiting order of the BatchNorm passes (engine.REV_*): blocks of steps per
variant, round robin.   python tools/order_ab.py [model]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import engine
from camvid_b200.nn import CrossEntropyLoss
from camvid_b200.optim import AdamW
from camvid_b200.utils import get_model

name = sys.argv[1] if len(sys.argv) > 1 else "unet"
variants = {"fff": (0, 0, 0), "rff": (1, 0, 0), "frf": (0, 1, 0), "rrf": (1, 1, 0), "rrr": (1, 1, 1), "ffr": (0, 0, 1), "frr": (0, 1, 1)}
torch.manual_seed(0)
net = get_model(name, 3, 12).cuda().train()
opt = AdamW(net.parameters(), lr=5e-4)
x = torch.randn(16, 3, 360, 480, device="cuda")
t = torch.randint(0, 12, (16, 360, 480), device="cuda")
loss_fn = CrossEntropyLoss()


def step():
    opt.zero_grad(set_to_none=True)
    loss_fn(net(x), t).backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
res = {k: [] for k in variants}
for rnd in range(5):
    for k, (a, r, b) in variants.items():
        engine.REV_APPLY, engine.REV_BWD_REDUCE, engine.REV_BWD_APPLY = bool(a), bool(r), bool(b)
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[k].append(e0.elapsed_time(e1) / 8)
base = sum(res["fff"]) / len(res["fff"])
print("variant = (forward apply, backward reduce, backward apply): f = front to back, r = reversed")
for k, v in res.items():
    m = sum(v) / len(v)
    print(f"{k}  mean {m:7.3f} ms  min {min(v):7.3f}  {m - base:+.3f} ms vs fff   {[round(a, 2) for a in v]}")
