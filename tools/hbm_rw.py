"""Read-only / write-only / copy bandwidth of this GPU's HBM with plain torch kernels (calibration for the memory-bound
kernels whose traffic is mostly stores or mostly loads).   python tools/hbm_rw.py"""
import torch

dev = torch.device("cuda")
n = 1 << 30  # bytes per buffer
a = torch.empty(n // 2, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)


def t(name, fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:28s} {best * 1e3:8.1f} us  {nbytes / best / 1e6:8.1f} GB/s")


t("write only (fill_)", lambda: b.fill_(1.0), n)
t("write only (cudaMemset)", lambda: b.view(torch.uint8).zero_(), n)
t("read only (sum fp32 view)", lambda: a.view(torch.float32).sum(), n)
t("copy (read + write)", lambda: b.copy_(a), 2 * n)
