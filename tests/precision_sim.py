"""CPU simulation: how much logit / gradient error do reduced-precision activation roundings cause in the reference
network at random init?  Straight-through rounding of conv operands (x, w) and conv outputs (y) inside the fp32 oracle.
    python tests/precision_sim.py unet 2 90 120"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import camvid_b200  # noqa
from camvid_b200.utils import get_model
from oracle import camvid_oracle as O
from util import rel_err


class Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt, gdt):
        ctx.gdt = gdt
        return x.to(dt).float() if dt is not None else x

    @staticmethod
    def backward(ctx, g):
        return (g.to(ctx.gdt).float() if ctx.gdt is not None else g), None, None


def make_cbr(op_dt, y_dt, g_dt):
    def cbr(x, sd, conv, bn, train, momentum=0.1, eps=1e-5):
        x = Round.apply(x, op_dt, g_dt)
        w = Round.apply(sd[conv + ".weight"], op_dt, None)
        y = F.conv2d(x, w, sd[conv + ".bias"], padding=1)
        y = Round.apply(y, y_dt, g_dt)
        y = F.batch_norm(y, sd[bn + ".running_mean"].clone(), sd[bn + ".running_var"].clone(), sd[bn + ".weight"],
                         sd[bn + ".bias"], train, momentum, eps)
        return F.relu(y)
    return cbr


name, n, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
torch.manual_seed(5)
sd = {k: v.clone() for k, v in get_model(name, 3, 12).state_dict().items()}
x, t = O.synth_batch(n, h, w, seed=6)
base = O.train_step(name, sd, x, t)
orig = O._cbr_impl
for tag, op_dt, y_dt, g_dt in [("bf16 ops, bf16 y, bf16 grads", torch.bfloat16, torch.bfloat16, torch.bfloat16),
                               ("bf16 ops, fp32 y, bf16 grads", torch.bfloat16, None, torch.bfloat16),
                               ("fp16 ops, fp16 y, fp32 grads", torch.float16, torch.float16, None),
                               ("fp16 ops, fp32 y, fp32 grads", torch.float16, None, None)]:
    O._cbr_impl = make_cbr(op_dt, y_dt, g_dt)
    loss, logits, grads, _ = O.train_step(name, sd, x, t)
    O._cbr_impl = orig
    agree = (logits.argmax(1) == base[1].argmax(1)).float().mean().item()
    ge = {k: rel_err(grads[k], base[2][k]) for k in grads if k.endswith("weight") and grads[k].dim() == 4}
    ks = list(ge)
    print(f"{tag}: logits rel {rel_err(logits, base[1]):.3e} loss rel {abs(loss - base[0]) / base[0]:.2e} argmax agree "
          f"{agree:.4f}; wgrad rel first {ge[ks[0]]:.2e} mid {ge[ks[len(ks) // 2]]:.2e} last {ge[ks[-1]]:.2e}")
