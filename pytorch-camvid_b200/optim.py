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
        if amsgrad or maximize or differentiable:
            raise ValueError("camvid_b200.optim.AdamW supports amsgrad=False, maximize=False, differentiable=False "
                             "(what train.py:100 uses)")
        if isinstance(lr, torch.Tensor):
            raise ValueError("camvid_b200.optim.AdamW takes a python float lr")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 \
                or not 0.0 <= weight_decay:
            raise ValueError(f"invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=bool(capturable), differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._tables = {}  # (group index, step) -> (key of device pointers, table tensor, chunk tensor)
        # capturable=True (CUDA graphs, camvid_b200.graph.GraphedTrainStep): the kernel reads its per-step scalar factors
        # (lr under OneCycleLR, bias corrections) from device memory; `advance()` recomputes them on the host and uploads
        # them stream-ordered, OUTSIDE the graph, before every replay.
        self._factors = {}  # group index -> device fp32 [8]

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
        # pinned + non_blocking: legal during CUDA-graph capture too (the gradients of a captured step live in the
        # graph's memory pool, so the table is rebuilt while capturing); the pinned sources are kept alive with the table
        host = (torch.tensor(rows, dtype=torch.int64).pin_memory(), torch.tensor(chunks, dtype=torch.int32).pin_memory())
        table, chunk_t = host[0].to(dev, non_blocking=True), host[1].to(dev, non_blocking=True)
        self._tables[gi] = (key, table, chunk_t, host)
        return table, chunk_t

    def _group_step(self, group):
        """The common step count of a group's parameters that have state (capturable mode keeps them in lockstep)."""
        steps = {int(self.state[p]["step"]) for p in group["params"] if len(self.state.get(p, {}))}
        if len(steps) > 1:
            raise RuntimeError("camvid_b200.optim.AdamW(capturable=True) needs one step count per param group")
        return steps.pop() if steps else 0

    def _upload_factors(self, gi, group, step):
        dev = next(p for p in group["params"] if p.grad is not None or len(self.state.get(p, {}))).device
        host = torch.empty(8, dtype=torch.float32, pin_memory=True)  # caching host allocator: safe to drop after the copy
        beta1, beta2 = group["betas"]
        rc = _lib.load().cvb_adamw_factors(float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                           float(group["weight_decay"]), step, ctypes.c_void_p(host.data_ptr()))
        if rc != 0:
            _lib.check(rc, "adamw_factors")
        with torch.cuda.device(dev):
            if gi not in self._factors:
                self._factors[gi] = torch.zeros(8, dtype=torch.float32, device=dev)
            self._factors[gi].copy_(host, non_blocking=True)

    @torch.no_grad()
    def advance(self):
        """capturable mode: what step() does on the HOST -- bump the step counts, recompute the scalar factors from the
        current lr / betas and upload them (stream-ordered copy) -- without launching the update. A captured graph that
        contains step() is replayed right after this."""
        for gi, group in enumerate(self.param_groups):
            if not group["capturable"]:
                raise RuntimeError("advance() is for AdamW(capturable=True)")
            step = self._group_step(group) + 1
            for p in group["params"]:
                if len(self.state.get(p, {})):
                    self.state[p]["step"].fill_(float(step))
            self._upload_factors(gi, group, step)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
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
                if not (capturing and group["capturable"]):  # a capture records the launch; advance() moves the counts
                    st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append(p)
            beta1, beta2 = group["betas"]
            if group["capturable"]:
                if len(by_step) > 1:
                    raise RuntimeError("camvid_b200.optim.AdamW(capturable=True) needs one step count per param group")
                for step, params in by_step.items():
                    if not capturing:
                        self._upload_factors(gi, group, step)
                    elif gi not in self._factors:
                        raise RuntimeError("capture a step only after at least one eager step() / advance()")
                    with torch.cuda.device(params[0].device):
                        table, chunks = self._table((gi, 0), step, params)
                        ops._call("adamw_step", 1, ("bytes", 28.0 * sum(p.numel() for p in params)),
                                  _lib.load().cvb_adamw_step_dev, ctypes.c_void_p(table.data_ptr()),
                                  ctypes.c_void_p(chunks.data_ptr()), chunks.shape[0],
                                  ctypes.c_void_p(self._factors[gi].data_ptr()), ops._stream())
                        torch.autograd.graph.increment_version(params)
                continue
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
