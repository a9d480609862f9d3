"""Experiment: the all-reduce kernel with world = 1 on plain device memory (bandwidth of the data loop alone)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import camvid_b200  # noqa
from camvid_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda")
n = int(138 * (1 << 20)) // 4
flat = torch.randn(n, device=dev)
ref = flat.clone()
flags = torch.zeros(lib.cvb_comm_flag_words(), dtype=torch.int32, device=dev)
bufs = (ctypes.c_void_p * 1)(flat.data_ptr())
fl = (ctypes.c_void_p * 1)(flags.data_ptr())
comm = _lib.Comm(ctypes.cast(bufs, ctypes.POINTER(ctypes.c_void_p)), ctypes.cast(fl, ctypes.POINTER(ctypes.c_void_p)), None, 0, 1)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
epoch = 0
for ctas in (16, 148, 592):
    ts = []
    for it in range(6):
        epoch += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.cvb_allreduce_mean_f32(ctypes.byref(comm), 0, n, 0, epoch, ctas, st)
        assert rc == 0, lib.cvb_last_error()
        e1.record()
        e2 = torch.cuda.Event(enable_timing=True)
        rc = lib.cvb_allreduce_wait(ctypes.byref(comm), 1, epoch, st)
        e2.record()
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    t = sorted(ts[2:])[1]
    print(f"ctas {ctas:4d}: kernel {t[0] * 1e3:8.1f} us ({2 * n * 4 / t[0] / 1e6:7.1f} GB/s r+w), wait kernel {t[1] * 1e3:6.1f} us, "
          f"unchanged: {torch.equal(flat, ref)}")
# ragged ranges: offsets / counts that are not multiples of 4 elements
for off, cnt in ((1, 1003), (3, 5), (2, 4097), (0, 7)):
    x = torch.randn(8192, device=dev)
    keep = x.clone()
    b2 = (ctypes.c_void_p * 1)(x.data_ptr())
    c2 = _lib.Comm(ctypes.cast(b2, ctypes.POINTER(ctypes.c_void_p)), ctypes.cast(fl, ctypes.POINTER(ctypes.c_void_p)), None, 0, 1)
    epoch += 1
    assert lib.cvb_allreduce_mean_f32(ctypes.byref(c2), off, cnt, 0, epoch, 4, st) == 0
    torch.cuda.synchronize()
    print("ragged", off, cnt, "unchanged:", torch.equal(x, keep))
