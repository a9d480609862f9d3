"""Micro-benchmark of the gradient all-reduce alone (no model): NCCL vs the library's NVLink kernel on a flat fp32 buffer.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_allreduce.py [MB ...]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import parallel

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
sizes = [float(a) for a in sys.argv[1:]] or [138.0, 25.0, 4.0]
for mb in sizes:
    n = int(mb * (1 << 20)) // 4
    cfgs = (("nccl", 0), ("nvlink", 16), ("nvlink", 32), ("nvlink", 64), ("nvlink", 148), ("nvls", 8), ("nvls", 16), ("nvls", 32),
            ("nvls", 64), ("nvls", 148))
    if os.environ.get("CVB_AR_DEBUG"):
        cfgs = (("nvlink", 148),)
    for backend, ctas in cfgs:
        if backend == "nccl":
            red = parallel.GradReducer(bucket_mb=1e9)
            flat = torch.empty(n, device=dev)
        else:
            red = parallel.PeerReducer(bucket_mb=1e9)
            red.CTAS = ctas
            red.NVLS = backend == "nvls"
            flat = red.buffer(n, dev)
        torch.manual_seed(rank)
        src = torch.randn(n, device=dev)
        ref = src.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        ts = []
        for it in range(8):
            flat.copy_(src)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            red.begin(flat)
            red.ready(0, n)
            red.finish()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        err = ((flat - ref).abs().max() / ref.abs().max()).item()
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        tm = torch.tensor([t], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{mb:7.1f} MB  {backend:6s} ctas {ctas:4d}  {tm.item() * 1e3:8.1f} us  algbw {mb * 1.048576 / tm.item():7.1f} GB/s  max err vs NCCL (rel. to max) {err:.2e}",
                  flush=True)
dist.destroy_process_group()
