"""Experiment: what bandwidth does a peer mapping of torch symmetric memory deliver to (a) the copy engine / torch
copy kernels and (b) plain torch elementwise kernels reading the peer tensor?  torchrun --nproc-per-node 2 ..."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 64 << 20  # 256 MB of fp32
buf = symm.empty(n, dtype=torch.float32, device=dev)
buf.fill_(float(rank + 1))
h = symm.rendezvous(buf, dist.group.WORLD)
peer = h.get_buffer((rank + 1) % world, (n,), torch.float32)
plain = torch.empty(n, device=dev)
torch.cuda.synchronize()
dist.barrier()


def timeit(name, fn, nbytes):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        print(f"{name:44s} {ms * 1e3:9.1f} us  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)
    dist.barrier()


nb = n * 4
timeit("local copy plain <- symm(local)", lambda: plain.copy_(buf), nb)
timeit("copy plain <- peer (read over NVLink)", lambda: plain.copy_(peer), nb)
timeit("copy peer <- plain (write over NVLink)", lambda: peer.copy_(plain), nb)
timeit("elementwise plain = peer * 2 (SM kernel, read)", lambda: torch.mul(peer, 2.0, out=plain), nb)
timeit("elementwise peer = plain * 2 (SM kernel, write)", lambda: torch.mul(plain, 2.0, out=peer), nb)
timeit("elementwise plain = peer + symm(local)", lambda: torch.add(peer, buf, out=plain), nb)
if rank == 0:
    print("has_multicast_support", h.has_multicast_support if hasattr(h, "has_multicast_support") else None,
          "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None, flush=True)
dist.destroy_process_group()
