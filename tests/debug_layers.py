"""Developer aid (not a test): per-block activation and per-parameter gradient errors of the CUDA plan vs the oracle.
    python tests/debug_layers.py unet 2 40 72 [synth|default]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import camvid_b200  # noqa
from camvid_b200.nn import CrossEntropyLoss
from camvid_b200.utils import get_model
from oracle import camvid_oracle as O
from util import rel_err

name, n, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
init = sys.argv[5] if len(sys.argv) > 5 else "synth"
torch.manual_seed(5)
net = get_model(name, 3, 12)
sd = O.synth_state_dict(net.state_dict(), seed=1) if init == "synth" else {k: v.clone() for k, v in net.state_dict().items()}
net.load_state_dict(sd)
net = net.cuda().train()
x, t = O.synth_batch(n, h, w, seed=2)
O.RECORD = {}
o_loss, o_logits, o_grads, _ = O.train_step(name, sd, x, t)
rec = O.RECORD
O.RECORD = None
logits = net(x.cuda())
loss = CrossEntropyLoss()(logits, t.cuda())
from camvid_b200 import engine
plan = engine.plans_of(net)[0]
torch.cuda.synchronize()
print("loss", loss.item(), o_loss.item(), "logits rel", rel_err(logits.cpu(), o_logits))
for b in plan.blocks:
    ref = rec[b.name + ".conv" if b.name.startswith("upsample") else b.name]["out"]
    got = b.a[..., :b.cout].float().permute(0, 3, 1, 2).cpu()
    if got.shape != ref.shape:  # UNet up-conv block writes a window of the concat buffer
        print(f"{b.name:28s} shape {tuple(got.shape)} vs {tuple(ref.shape)}")
        continue
    print(f"{b.name:28s} act rel {rel_err(got, ref):.3e}  x{tuple(b.x.shape)} -> a{tuple(b.a.shape)} taps {b.taps}")
loss.backward()
torch.cuda.synchronize()
for k, p in net.named_parameters():
    g = p.grad.cpu()
    print(f"{k:40s} grad rel {rel_err(g, o_grads[k]):.3e}  |g| {g.norm():.3e} |ref| {o_grads[k].norm():.3e}")
