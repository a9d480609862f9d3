"""Whole-step CUDA graph: zero_grad -> net(x) -> loss -> backward -> optimizer.step() (train.py:124-134) captured once and
replayed, so that a training iteration costs the host one graph launch instead of ~260 kernel launches through the C ABI,
and dependent kernels follow each other without launch gaps.

    step = GraphedTrainStep(net, camvid_b200.nn.CrossEntropyLoss(), camvid_b200.optim.AdamW(net.parameters(), lr,
                                                                                          capturable=True), x0, t0)
    for images, masks in prefetcher:
        loss = step(images, masks)          # device scalar; scheduler.step() etc. as usual afterwards

What makes the path capturable: every buffer of a forward / backward pass belongs to the execution plan (engine.Plan) or
is allocated from the graph's private pool while capturing; the weight-gradient side stream forks from and joins the
capturing stream through events; the optimizer's per-step scalars (lr under OneCycleLR, bias corrections) are read from
device memory and refreshed by `optimizer.advance()` right before each replay; the bf16 GEMM operands are re-packed
from the fp32 weights inside the graph.

Not captured (stays eager): data-parallel runs (the NCCL reducer), eval-mode forwards.

Build it while no earlier loss / output of the network is alive: an old autograd graph keeps the parameters'
AccumulateGrad nodes bound to the stream it ran on, and a capture cannot make that stream wait on the capturing one
(torch's own rule for whole-network capture; the constructor turns the CUDA error into this hint).
"""
import torch

from . import engine


class GraphedTrainStep:
    def __init__(self, net, loss_fn, optimizer, example_images, example_masks, warmup=2):
        if not getattr(optimizer, "defaults", {}).get("capturable", False) or not hasattr(optimizer, "advance"):
            raise RuntimeError("GraphedTrainStep needs camvid_b200.optim.AdamW(..., capturable=True)")
        if net.__dict__.get("_cvb_reducer") is not None:
            raise RuntimeError("GraphedTrainStep does not capture data-parallel steps (the NCCL reducer stays eager)")
        if not example_images.is_cuda:
            raise RuntimeError("camvid_b200 runs on CUDA only; there is no CPU path")
        self.net, self.loss_fn, self.optimizer = net, loss_fn, optimizer
        self.images = example_images.detach().clone()
        self.masks = example_masks.detach().clone()
        self.params = [p for p in net.parameters()]
        dev = self.images.device
        net.train()
        with torch.cuda.device(dev):
            # Warm-up on a side stream (plans, packed operands, optimizer state and tables get built; lazy CUDA
            # initialisation happens outside the capture). The warm-up steps are REAL steps on the example batch, so
            # everything they touch is put back afterwards: constructing the graph has no effect on training.
            saved = self._snapshot()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._eager_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._restore(saved)
            optimizer.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self.graph):
                    self.loss = self._eager_step()
            except RuntimeError as e:
                if "capturing" in str(e) or "capture" in str(e):
                    raise RuntimeError("GraphedTrainStep: the capture was invalidated. The usual cause is a loss / logits "
                                       "tensor of an earlier eager step that is still alive (its autograd graph pins the "
                                       "parameters' AccumulateGrad nodes to the default stream): `del loss` before "
                                       "building the graph. Original error: " + str(e)) from e
                raise
        self.launches_per_replay = None

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.net(self.images), self.masks)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _snapshot(self):
        state = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.net.state_dict().items()}
        opt = {id(p): {k: v.clone() for k, v in st.items()} for p, st in self.optimizer.state.items()}
        return state, opt

    def _restore(self, saved):
        state, opt = saved
        with torch.no_grad():
            for k, v in self.net.state_dict().items():
                v.copy_(state[k])
            for p, st in self.optimizer.state.items():
                old = opt.get(id(p))
                for k, v in st.items():
                    if old is None:  # state created by the warm-up: back to "never stepped"
                        v.zero_()
                    else:
                        v.copy_(old[k])
        torch.autograd.graph.increment_version(self.params)

    def __call__(self, images, masks):
        """One training step on this batch (same shapes / dtypes as the example batch). Returns the loss as a device
        scalar that the NEXT call overwrites."""
        if images.shape != self.images.shape or masks.shape != self.masks.shape or masks.dtype != self.masks.dtype:
            raise RuntimeError("GraphedTrainStep replays one input geometry; build another instance for "
                               f"{tuple(images.shape)} / {tuple(masks.shape)} {masks.dtype}")
        with torch.cuda.device(self.images.device):
            self.images.copy_(images, non_blocking=True)
            self.masks.copy_(masks, non_blocking=True)
            self.optimizer.advance()  # host: step counts, lr / bias-correction factors -> device, ordered before the replay
            self.graph.replay()
        # the replay re-packed the GEMM operands from the weights as they were BEFORE its own optimizer update, and wrote
        # parameters and running statistics through raw pointers: tell torch, so that the next eager forward re-packs
        torch.autograd.graph.increment_version(self.params)
        engine.touch_running_stats(self.net)
        return self.loss
