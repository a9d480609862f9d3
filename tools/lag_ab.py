"""Interleaved in-process comparison of weight-gradient lags (engine.WGRAD_LAG): blocks of steps per variant, round
robin, so that the box's clock drift hits every variant alike.   python tools/lag_ab.py [model] [lags...]"""
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
lags = [int(a) for a in sys.argv[2:]] or [0, 2, 4, 6, 8, 12]
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
res = {lag: [] for lag in lags}
for rnd in range(5):
    for lag in lags:
        engine.WGRAD_LAG = lag
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[lag].append(e0.elapsed_time(e1) / 8)
base = sum(res[lags[0]]) / len(res[lags[0]])
for lag in lags:
    v = res[lag]
    m = sum(v) / len(v)
    print(f"lag {lag:3d}  mean {m:7.3f} ms  min {min(v):7.3f}  {m - base:+.3f} ms vs lag {lags[0]}   {[round(a, 2) for a in v]}")
