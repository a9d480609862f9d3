"""Interleaved comparison of gradient-exchange variants inside ONE process per GPU (the step runs at the power cap and the
clocks drift by several percent over a minute: separate runs cannot resolve sub-millisecond differences).
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_ab.py [rounds] [steps_per_block]
Every round runs each variant for `steps_per_block` steps, in rotating order; rank 0 prints the mean / min step time per
variant over the rounds and the difference to "no exchange"."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import parallel
from camvid_b200.nn import CrossEntropyLoss
from camvid_b200.optim import AdamW
from camvid_b200.utils import get_model

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
block = int(sys.argv[2]) if len(sys.argv) > 2 else 8
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
B, H, W = 16, 360, 480
torch.manual_seed(0)
net = get_model("unet", 3, 12).to(dev).train()
parallel.data_parallel(net)  # broadcast + default reducer (replaced below)
loss_fn = CrossEntropyLoss()
opt = AdamW(net.parameters(), lr=5e-4, weight_decay=0)
g = torch.Generator().manual_seed(1 + rank)
xs = [torch.randn(B, 3, H, W, generator=g).to(dev) for _ in range(2)]
ts = [torch.randint(0, 12, (B, H, W), generator=g).to(dev) for _ in range(2)]


def nccl(bucket_mb):
    return parallel.GradReducer(bucket_mb=bucket_mb)


def nvlink(bucket_mb, ctas, nvls=False):
    r = parallel.PeerReducer(bucket_mb=bucket_mb)
    r.CTAS = ctas
    r.NVLS = nvls
    return r


variants = {
    "none": None,
    "nccl_25MB": nccl(25),
    "nvlink_25MB_32": nvlink(25, 32),
    "nvlink_8MB_32": nvlink(8, 32),
    "nvlink_8MB_64": nvlink(8, 64),
    "nvls_8MB_16": nvlink(8, 16, True),
    "nvls_8MB_32": nvlink(8, 32, True),
    "nvls_25MB_16": nvlink(25, 16, True),
    "nvls_4MB_16": nvlink(4, 16, True),
}
names = list(variants)


def step(i):
    opt.zero_grad(set_to_none=True)
    loss = loss_fn(net(xs[i % 2]), ts[i % 2])
    loss.backward()
    opt.step()


def run_block(name, n):
    r = variants[name]
    if r is None:
        net.__dict__.pop("_cvb_reducer", None)
    else:
        net.__dict__["_cvb_reducer"] = r
    step(0)  # switch-over step (buffers, tables), untimed
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for name in names:  # warm every variant once
    run_block(name, 2)
res = {n: [] for n in names}
for rd in range(rounds):
    order = names[rd % len(names):] + names[:rd % len(names)]
    for name in order:
        res[name].append(run_block(name, block))
if rank == 0:
    base = sum(res["none"]) / len(res["none"])
    out = {}
    for n in names:
        m = sum(res[n]) / len(res[n])
        out[n] = {"mean_ms": round(m, 3), "min_ms": round(min(res[n]), 3), "delta_vs_none_ms": round(m - base, 3),
                  "samples": [round(v, 2) for v in res[n]]}
        print(f"{n:18s} mean {m:7.3f} ms  min {min(res[n]):7.3f}  +{m - base:6.3f} ms vs none   {out[n]['samples']}", flush=True)
    print(json.dumps({"world": world, "rounds": rounds, "steps_per_block": block, "variants": out}))
dist.destroy_process_group()
