"""Drop-in for the optimizer the reference trains with: `optim.AdamW(net.parameters(), lr=args.lr, weight_decay=args.wd)`
(train.py:100) stepped once per iteration (train.py:133) under `OneCycleLR` (train.py:102-104, which rewrites `lr` and
`betas` of the param groups every step). Same constructor, same `param_groups` / `state` layout (`step`, `exp_avg`,
`exp_avg_sq`: state dicts are interchangeable with torch.optim.AdamW), one fused kernel launch per param group instead of
the ~8 multi-tensor passes of the stock implementation.
"""
import ctypes

import torch

from . import _lib, ops


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None):
        if amsgrad or maximize or capturable or differentiable:
            raise ValueError("camvid_b200.optim.AdamW supports amsgrad=False, maximize=False, capturable=False, "
                             "differentiable=False (what train.py:100 uses)")
        if isinstance(lr, torch.Tensor):
            raise ValueError("camvid_b200.optim.AdamW takes a python float lr")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 \
                or not 0.0 <= weight_decay:
            raise ValueError(f"invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._tables = {}  # (group index, step) -> (key of device pointers, table tensor, chunk tensor)

    def _table(self, gi, step, params):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr(), p.numel()) for p in params)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        chunk = _lib.load().cvb_adamw_chunk_elems()
        rows, chunks = [], []
        for i, k in enumerate(key):
            rows.append(list(k))
            chunks += [[i, c] for c in range((k[4] + chunk - 1) // chunk)]
        dev = params[0].device
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        chunk_t = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self._tables[gi] = (key, table, chunk_t)
        return table, chunk_t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g.dtype == torch.float32
                        and g.is_contiguous() and not g.is_sparse and g.device == p.device):
                    raise RuntimeError("camvid_b200.optim.AdamW updates contiguous fp32 CUDA parameters with dense fp32 "
                                       "gradients only (there is no CPU path)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append(p)
            beta1, beta2 = group["betas"]
            for k, (step, params) in enumerate(sorted(by_step.items())):
                with torch.cuda.device(params[0].device):
                    table, chunks = self._table((gi, k), step, params)
                    ops._call("adamw_step", 1, ("bytes", 28.0 * sum(p.numel() for p in params)),
                              _lib.load().cvb_adamw_step, ctypes.c_void_p(table.data_ptr()),
                              ctypes.c_void_p(chunks.data_ptr()), chunks.shape[0], float(group["lr"]), float(beta1),
                              float(beta2), float(group["eps"]), float(group["weight_decay"]), step, ops._stream())
                # The kernel wrote the parameters through raw pointers: tell torch, exactly as an in-place op would.
                # The execution plans key their packed bf16 GEMM operands on (data_ptr, _version) of the fp32 weight
                # (engine.Block._weight_key) and re-pack when it moves; without this they would keep convolving with
                # the initial weights.
                torch.autograd.graph.increment_version(params)
        return loss
