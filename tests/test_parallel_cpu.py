"""CPU tests (gloo, world_size 2) of the data-parallel host logic: the bucketed gradient reducer that the execution
plans drive during backward (camvid_b200/parallel.py), and the confusion-matrix reduction used by eval."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import camvid_b200  # noqa: F401
from camvid_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        base = torch.randn(1000)
        flat = base * (rank + 1)  # rank r holds (r+1) * base: the mean over 2 ranks is 1.5 * base
        red = parallel.GradReducer(bucket_mb=4 * 300 / (1 << 20))  # 300-element buckets
        red.begin(flat)
        launched = []
        for lo, hi in ((0, 120), (120, 310), (310, 330), (330, 900), (900, 1000)):  # ranges as blocks finish
            red.ready(lo, hi)
            launched.append(red.buckets_launched)
        red.finish()
        ok = torch.allclose(flat, 1.5 * base, rtol=1e-6, atol=1e-6)
        # buckets are cut when >= 300 elements are pending: after 310, after 900, and the tail at finish()
        ok = ok and launched == [0, 1, 1, 2, 2] and red.buckets_launched == 3
        # out-of-order ranges are rejected loudly
        red.begin(torch.zeros(10))
        try:
            red.ready(2, 5)
            ok = False
        except RuntimeError:
            pass
        # a marked module broadcasts rank 0's parameters
        lin = torch.nn.Linear(3, 2)
        with torch.no_grad():
            lin.weight.fill_(float(rank + 1))
        parallel.data_parallel(lin)
        ok = ok and bool((lin.weight == 1.0).all()) and "_cvb_reducer" in lin.__dict__
        cm = torch.full((12, 12), rank + 1, dtype=torch.int64)
        parallel.all_reduce_confusion(cm)
        ok = ok and bool((cm == 3).all())
        result[rank] = ok
    finally:
        dist.destroy_process_group()


def test_grad_reducer_and_confusion_allreduce_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        result = mgr.dict()
        mp.spawn(_worker, args=(world, port, result), nprocs=world, join=True)
        assert dict(result) == {0: True, 1: True}


def test_reducer_requires_initialised_process_group():
    if dist.is_initialized():
        pytest.skip("a process group is active in this process")
    with pytest.raises(RuntimeError, match="not initialised"):
        parallel.GradReducer()
