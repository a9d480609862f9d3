#!/usr/bin/env python
"""Benchmark of the hot path: UNet (or SegNet) training step on CamVid-shaped synthetic batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model unet|segnet]
                    [--batch 16] [--height 360] [--width 480]

A step = what train.py:124-134 does per iteration: zero_grad, forward, CrossEntropyLoss, backward, AdamW.step.
N > 1 is launched by torchrun (one rank per GPU); per-GPU batch is fixed (weak scaling), gradients are all-reduced
over NCCL in buckets on a side stream. Rank 0 prints ONE JSON line.

  value         images/s, whole job, inputs resident in HBM, CUDA-event time over exactly K steps, max over ranks
  e2e           the same through the public API with HOST inputs: camvid_b200.data.DevicePrefetcher copies each step's uint8
                HWC images + uint8 masks (what cv2 / the dataset yield) from pageable host memory through its pinned ring
                on a side stream and runs ToTensor + Normalize on the device; the loss is read back every step
  roofline      the dominant kernel family (cvb_conv3x3_fprop: every forward conv and every data-gradient conv):
                algorithmic FLOPs of its launches / their CUDA-event durations, measured on K further steps of the same
                workload with an event pair around every C-ABI call (kept out of `value` so the events cannot perturb it)
  kernels       the same arithmetic for every other kernel family (HBM-bound ones in GB/s)
  cpu_baseline  the reference's own training step (oracle/_ref: its unmodified modules, staged by oracle/make_ref.py)
                timed on this box's host cores on a bounded sample; the oracle port if oracle/_ref is absent
  dp_check      N > 1: gradients of one batch through the bucketed NCCL reducer against the all-reduced mean of the
                rank-local gradients, and the spread of the parameters over the ranks after the timed steps

`--impl reference` times the reference's own CPU path (oracle/_ref, else the oracle port) with all host threads on
BASELINE.md section 5's fixed sample (batch 2 of the same geometry), same metric / unit.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "images/s"


def metric_name(args):
    return f"{'UNet' if args.model == 'unet' else 'SegNet'} train images/sec @3x{args.height}x{args.width} bf16"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="unet", choices=["unet", "segnet"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--mode", default="train", choices=["train", "eval"],
                    help="eval: BASELINE config 5 (forward with running statistics + argmax + confusion matrix / mIoU)")
    ap.add_argument("--optimizer", default="b200", choices=["b200", "torch"],
                    help="b200 = camvid_b200.optim.AdamW (fused drop-in, parity-tested against torch); torch = torch.optim.AdamW")
    ap.add_argument("--graph", action="store_true",
                    help="also time the step as ONE CUDA graph replay (camvid_b200.graph.GraphedTrainStep; single GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=2)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained",
                                                                                    p["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def traffic_from_profiles(kernel):
    """DRAM bytes per launch of a kernel family from the committed ncu capture (profiles/traffic.json), else None.
    Only meaningful for the workload the capture was taken on (UNet 16x3x360x480)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return round(json.load(open(path))[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_run(model, sample_batch, h, w, steps, warmup, threads=None):
    """The reference's training step (get_model + CrossEntropyLoss + AdamW, fp32, train.py:100-134) on the host cores:
    the reference's own modules when oracle/_ref is staged (kind "reference"), else the oracle port (kind "port").
    Returns (images/s, seconds per step, threads, kind)."""
    import torch
    from oracle import ref_runner
    if ref_runner.available():
        v, sec, thr, _ = ref_runner.train_steps(model, sample_batch, h, w, steps, warmup, threads)
        return v, sec, thr, "reference"
    from oracle import camvid_oracle as O
    import camvid_b200  # noqa: F401  (module tree only: parameter names / shapes / default init; never run on CPU)
    from camvid_b200.utils import get_model
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in get_model(model, 3, 12).state_dict().items()}
    names = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    params = [torch.nn.Parameter(sd[k].clone()) for k in names]
    opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0)
    x, t = O.synth_batch(sample_batch, h, w, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        for k, p in zip(names, params):
            sd[k] = p.detach()
        loss, _, grads, after = O.train_step(model, sd, x, t)
        for k, p in zip(names, params):
            p.grad = grads[k]
        opt.step()
        sd = after
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return sample_batch / sec, sec, torch.get_num_threads(), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # a fixed batch-2 sample of the workload's geometry per step; the number of timed steps is cut
    # (never below 5) so that the whole run stays near 3 minutes
    sb = args.cpu_sample_batch
    _, probe, _, _ = cpu_reference_run(args.model, sb, args.height, args.width, 1, 0, cores)
    steps = max(5, min(args.steps, int(150.0 / probe) - args.warmup))
    warm = max(1, min(args.warmup, 3))
    val, sec, thr, kind = cpu_reference_run(args.model, sb, args.height, args.width, steps, warm, cores)
    sample = (f"{steps} timed steps (+{warm} warm-up) of {args.model} fwd+loss+bwd+AdamW fp32 on a {sb}x3x{args.height}x"
              f"{args.width} sample of the {args.batch}x3x{args.height}x{args.width} per-GPU batch, {sec:.2f} s/step, "
              + ("the reference's own modules (oracle/_ref)" if kind == "reference" else "oracle port (oracle/_ref absent)"))
    cfg = dict(workload_config(args), sample=sample, reference_sample_batch=sb, optimizer="torch.optim.AdamW")
    out = {"impl": "reference", "metric": metric_name(args),
           "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": thr, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(args):
    return {"workload": f"{args.model} train step (fwd + CrossEntropyLoss + bwd + AdamW), per-GPU batch "
                        f"{args.batch}x3x{args.height}x{args.width}, 12 classes, random init",
            "model_family": args.model, "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus,
            "height": args.height, "width": args.width, "classes": 12,
            "parallelism": f"dp{args.gpus}" if args.gpus > 1 else "single",
            "optimizer": "camvid_b200.optim.AdamW (fused)" if getattr(args, "optimizer", "b200") == "b200" else "torch.optim.AdamW",
            "l2": "no flush needed: one step streams several GB of activations, far beyond the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import camvid_b200  # noqa: F401
    from camvid_b200 import ops, parallel
    from camvid_b200.nn import CrossEntropyLoss
    from camvid_b200.utils import get_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The first communicator prints an "NCCL version ..." banner on stdout (from native code): point fd 1 at stderr
        # while the communicator is created, stdout carries the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if args.gpus != world and rank == 0:
        print(f"note: --gpus {args.gpus} but WORLD_SIZE {world}; using {world}", file=sys.stderr)
    args.gpus = world

    B, H, W = args.batch, args.height, args.width
    torch.manual_seed(0)
    net = get_model(args.model, 3, 12).to(dev).train()
    if world > 1:
        parallel.data_parallel(net)
    loss_fn = CrossEntropyLoss()
    if args.optimizer == "b200":
        from camvid_b200.optim import AdamW
    else:
        AdamW = torch.optim.AdamW
    opt = AdamW(net.parameters(), lr=5e-4, weight_decay=0)  # train.py:100

    g = torch.Generator().manual_seed(1 + rank)
    nbuf = 2
    host_x = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(nbuf)]
    host_t = [torch.randint(0, 12, (B, H, W), generator=g).pin_memory() for _ in range(nbuf)]
    dev_x = [x.to(dev) for x in host_x]
    dev_t = [t.to(dev) for t in host_t]

    # HOST batches as cv2 / the dataset yield them (dataset/camvid.py:161-173): uint8 HWC images + uint8 masks, pageable
    from camvid_b200 import data
    host_u8 = [torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8) for _ in range(nbuf)]
    host_m8 = [t.to(torch.uint8) for t in host_t]

    class Loader:
        def __init__(self, k):
            self.k = k

        def __len__(self):
            return self.k

        def __iter__(self):
            for i in range(self.k):
                yield host_u8[i % nbuf], host_m8[i % nbuf]

    def step(x, t):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net(x), t)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tns = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return tns.item()

    if args.mode == "eval":
        run_eval(args, net, dev_x, dev_t, Loader, barrier, max_over_ranks, world, rank, dev)
        if world > 1:
            dist.destroy_process_group()
        return

    for i in range(max(args.warmup, 3)):
        step(dev_x[i % nbuf], dev_t[i % nbuf])
    barrier()

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(dev_x[i % nbuf], dev_t[i % nbuf])
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps  # host time to ENQUEUE a step (the GPU runs behind)
    e1.record()
    barrier()
    clocks = sampler.result()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = (ops.LAUNCHES - l0) // args.steps
    last_loss = loss.item()
    del loss  # (an old loss keeps its autograd graph, and with it AccumulateGrad nodes bound to this stream, alive)
    value = world * B * args.steps / (ms * 1e-3)

    # ---- the same step as one CUDA graph replay (optional)
    graph_info = None
    if args.graph and world == 1 and args.optimizer == "b200":
        from camvid_b200.graph import GraphedTrainStep
        gopt = AdamW(net.parameters(), lr=5e-4, weight_decay=0, capturable=True)
        l_before = ops.LAUNCHES
        gstep = GraphedTrainStep(net, loss_fn, gopt, dev_x[0], dev_t[0], warmup=2)
        captured = (ops.LAUNCHES - l_before) // 3  # two warm-up steps + the captured one
        for i in range(3):
            gstep(dev_x[i % nbuf], dev_t[i % nbuf])
        barrier()
        e0.record()
        h0 = time.perf_counter()
        for i in range(args.steps):
            gloss = gstep(dev_x[i % nbuf], dev_t[i % nbuf])
        ghost_ms = (time.perf_counter() - h0) * 1e3 / args.steps
        e1.record()
        barrier()
        gms = e0.elapsed_time(e1) / args.steps
        graph_info = {"value": B / (gms * 1e-3), "unit": UNIT, "ms_per_step": gms, "host_ms_per_step": round(ghost_ms, 3),
                      "kernels_in_graph": int(captured), "loss": gloss.item(),
                      "what": "zero_grad + forward + loss + backward + AdamW captured once (GraphedTrainStep), replayed per step; "
                              "inputs copied device-to-device into the graph's static buffers inside the timed region"}

    # ---- end to end: HOST inputs through the library's own input stage (camvid_b200.data.DevicePrefetcher): uint8 HWC
    # images + uint8 masks in pageable memory, as cv2 / the dataset yield them -> pinned ring -> side-stream H2D ->
    # ToTensor + Normalize on the device (cvb_input_stage_u8) -> fp32 NCHW images + uint8 masks for net / loss. The
    # loss of step i is read on the host after step i+1 has been enqueued (one device->host read per step without
    # draining the queue), like a training script that prints the previous iteration's loss.
    # One GPU: the step itself is the library's CUDA-graph step (GraphedTrainStep: same kernels, same bits, 0.3 ms of host
    # time per step instead of 15), so that the host has room for the input pipeline; data parallel runs stay eager.
    e2e_step, e2e_mode = step, "eager step"
    if world == 1 and args.optimizer == "b200":
        try:
            from camvid_b200.graph import GraphedTrainStep
            if graph_info is None:
                gopt = AdamW(net.parameters(), lr=5e-4, weight_decay=0, capturable=True)
            with torch.no_grad():
                x0 = torch.empty(B, 3, H, W, device=dev).normal_()
                m0 = host_m8[0].to(dev)
            gstep8 = GraphedTrainStep(net, loss_fn, gopt, x0, m0, warmup=1)
            e2e_step, e2e_mode = gstep8, "CUDA-graph step (camvid_b200.graph.GraphedTrainStep)"
        except Exception as exc:  # never lose the bench line over the optional path
            print(f"note: graphed e2e step unavailable ({type(exc).__name__}: {exc}); using the eager step", file=sys.stderr)

    def e2e_loop(k):
        pf = data.DevicePrefetcher(Loader(k), dev, mask_dtype=torch.uint8)
        pending = None
        for x, m in pf:
            loss_i = e2e_step(x, m)
            host_loss = torch.empty((), dtype=torch.float32, pin_memory=True)
            host_loss.copy_(loss_i.detach(), non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            if pending is not None:
                pending[1].synchronize()
                float(pending[0])
            pending = (host_loss, done)
        pending[1].synchronize()
        return float(pending[0]), pf.h2d_bytes

    e2e_loop(3)
    barrier()
    e0.record()
    w0 = time.perf_counter()
    _, h2d = e2e_loop(args.steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
    e2e = {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": world * h2d // args.steps, "d2h_bytes_per_step": 4 * world,
           "ms_per_step": e2e_ms / args.steps,
           "how": "camvid_b200.data.DevicePrefetcher: uint8 HWC images + uint8 masks from pageable host memory, pinned "
                  "ring, side-stream copy, ToTensor + Normalize on the device; " + e2e_mode + "; loss read back with a "
                  "one-step lag"}

    # ---- data-parallel sanity (N > 1): the bucketed NCCL reducer against a plain all-reduce of rank-local gradients
    dp_check = None
    if world > 1:
        def flat_grads():
            return torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        x0, t0 = dev_x[0], dev_t[0]
        opt.zero_grad(set_to_none=True)
        loss_fn(net(x0), t0).backward()
        g_dp = flat_grads().clone()
        reducer = net.__dict__.pop("_cvb_reducer")
        opt.zero_grad(set_to_none=True)
        loss_fn(net(x0), t0).backward()
        g_ref = flat_grads().clone()
        net.__dict__["_cvb_reducer"] = reducer
        g_local_norm = g_ref.norm().item()
        dist.all_reduce(g_ref, op=dist.ReduceOp.AVG)
        pf_ = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        pmax, pmin = pf_.clone(), pf_.clone()
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        dp_check = {"transport": type(reducer).__name__ + (" (NVLS)" if getattr(reducer, "multicast", False) else ""),
                    "grad_rel_err": ((g_dp - g_ref).norm() / g_ref.norm()).item(),
                    "grad_max_abs_err": (g_dp - g_ref).abs().max().item(),
                    "mean_grad_norm": g_ref.norm().item(), "rank0_local_grad_norm": g_local_norm,
                    "param_spread_over_ranks_max_abs": (pmax - pmin).abs().max().item(),
                    "buckets": reducer.buckets_launched,
                    "what": "gradients of one batch per rank through the bucketed side-stream reducer vs all_reduce(AVG) of "
                            "the rank-local gradients of the same batches; max |param(rank i) - param(rank j)| after all steps"}
        opt.zero_grad(set_to_none=True)

    # ---- per-kernel CUDA-event timing (same workload, K further steps)
    pk = peaks()
    roofline, roofline_wgrad, kernels, diagnostics = None, None, {}, None
    if not args.no_kernel_timing:
        from camvid_b200 import engine
        engine.OVERLAP_WGRAD = False  # one stream: every event pair brackets exactly one kernel
        ksteps = min(args.steps, 10)
        # diagnostic: the same steps on ONE stream without the event pairs -- minus the sum of the kernel times below, this
        # is what the gaps between dependent launches cost
        for i in range(2):
            step(dev_x[i % nbuf], dev_t[i % nbuf])
        barrier()
        e0.record()
        for i in range(ksteps):
            step(dev_x[i % nbuf], dev_t[i % nbuf])
        e1.record()
        barrier()
        serial_ms = e0.elapsed_time(e1) / ksteps
        rec = ops.profile(True)
        barrier()
        for i in range(ksteps):
            step(dev_x[i % nbuf], dev_t[i % nbuf])
        barrier()
        ops.profile(False)
        engine.OVERLAP_WGRAD = True
        agg = {}
        algo_bytes = {}
        for what, work, a, b in rec:
            d = agg.setdefault(what, [0.0, 0.0, 0, work[0]])
            d[0] += a.elapsed_time(b) * 1e-3
            d[1] += work[1]
            d[2] += 1
            if len(work) > 3:
                algo_bytes[what] = algo_bytes.get(what, 0.0) + work[3]
        step_s = sum(d[0] for d in agg.values()) / ksteps
        diagnostics = {"single_stream_ms_per_step": round(serial_ms, 3), "kernel_time_sum_ms_per_step": round(step_s * 1e3, 3),
                       "launch_gap_ms_per_step": round(serial_ms - step_s * 1e3, 3),
                       "what": "one stream, weight gradients not overlapped: step time vs the sum of its kernels' durations"}
        for what, (sec, amount, n, kind) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            if kind == "flops":
                ach, peak, unit, bound = amount / sec / 1e12, pk["tf_sustained"], "TFLOP/s", "tensor"
            else:
                ach, peak, unit, bound = amount / sec / 1e9, pk["hbm"], "GB/s", "hbm"
            kernels[what] = {"bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit,
                             "frac": round(ach / peak, 4), "launches_per_step": n // ksteps,
                             "avg_us": round(sec / n * 1e6, 2), "share_of_step": round(sec / ksteps / step_s, 4)}
        top = kernels.get("conv3x3_fprop")
        if top:
            roofline = {"kernel": "conv3x3 forward + data-gradient family (conv_fprop_kernel<256>, conv_fprop_tr128_kernel, conv_fprop_tr64_kernel)", "bound": "tensor",
                        "achieved": top["achieved"], "peak": top["peak"], "unit": "TFLOP/s", "frac": top["frac"],
                        "peak_source": f"{pk['src']} sustained cuBLAS bf16 (kernel timed inside a long step)",
                        "frac_of_burst_peak": round(top["achieved"] / pk["tf_burst"], 4),
                        "traffic": traffic_from_profiles("conv3x3_fprop"),
                        "algorithmic_bytes_per_launch": round(algo_bytes.get("conv3x3_fprop", 0.0) / max(agg["conv3x3_fprop"][2], 1)),
                        "avg_launch_us": top["avg_us"], "launches_per_step": top["launches_per_step"],
                        "share_of_step": top["share_of_step"]}
        wg = kernels.get("conv3x3_wgrad")
        if wg:
            roofline_wgrad = {"kernel": "conv3x3 weight-gradient family (conv_wgrad_kernel<BN>, conv_wgrad_rs64_kernel, + split-K "
                                        "part-sum / OIHW transpose)", "bound": "tensor", "achieved": wg["achieved"],
                              "peak": wg["peak"], "unit": "TFLOP/s", "frac": wg["frac"],
                              "frac_of_burst_peak": round(wg["achieved"] / pk["tf_burst"], 4),
                              "traffic": traffic_from_profiles("conv3x3_wgrad"), "avg_launch_us": wg["avg_us"],
                              "launches_per_step": wg["launches_per_step"], "share_of_step": wg["share_of_step"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    from camvid_b200 import engine as _engine
    plan = _engine.plans_of(net)[0]
    flops_img = sum(b.flops * (3 if i > 0 else 2) for i, b in enumerate(plan.blocks)) / B
    out = {"metric": metric_name(args), "value": value,
           "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic", "config": workload_config(args), "e2e": e2e,
           "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
           "host_enqueue_ms_per_step": round(host_ms, 3), "graph_replay": graph_info, "clocks": clocks,
           "loss": last_loss, "conv_gflop_per_image": round(flops_img / 1e9, 2),
           "conv_tensor_util_end_to_end": round(value / world * flops_img / 1e12 / pk["tf_sustained"], 4),
           "roofline": roofline, "roofline_wgrad": roofline_wgrad, "kernels": kernels, "diagnostics": diagnostics}
    if dp_check is not None:
        out["dp_check"] = dp_check
    if world == 1 and not args.no_cpu_baseline:
        sb = args.cpu_sample_batch
        v, sec, thr, kind = cpu_reference_run(args.model, sb, H, W, 5, 1)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": thr, "kind": kind,
                               "sample": f"5 timed steps (+1 warm-up) of the reference train step (fwd + loss + bwd + "
                                         f"AdamW, fp32) on a {sb}x3x{H}x{W} batch, {sec:.2f} s/step, "
                                         + ("the reference's own modules (oracle/_ref)" if kind == "reference"
                                            else "oracle port (oracle/_ref absent)")}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_eval(args, net, dev_x, dev_t, Loader, barrier, max_over_ranks, world, rank, dev):
    """eval.py:50-72 / train.py:180-197 on the device: forward with BatchNorm folded into the conv epilogues, fused
    argmax + confusion matrix, mIoU from the (all-reduced) matrix."""
    import torch
    from camvid_b200 import data, ops, parallel
    from camvid_b200.legacy.metrics import Metrics
    net.eval()
    B, nbuf = args.batch, len(dev_x)
    cm = torch.zeros(12, 12, dtype=torch.int64, device=dev)

    def estep(x, t):
        with torch.no_grad():
            logits = net(x)
            ops.argmax_confusion_nchw(logits, t, cm)

    for i in range(max(args.warmup, 3)):
        estep(dev_x[i % nbuf], dev_t[i % nbuf])
    barrier()
    cm.zero_()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        estep(dev_x[i % nbuf], dev_t[i % nbuf])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = (ops.LAUNCHES - l0) // args.steps
    pf = data.DevicePrefetcher(Loader(3), dev, mask_dtype=torch.uint8)
    for x, m in pf:  # warm the ring
        estep(x, m)
    barrier()
    cm.zero_()
    e0.record()
    w0 = time.perf_counter()
    pf = data.DevicePrefetcher(Loader(args.steps), dev, mask_dtype=torch.uint8)
    for x, m in pf:  # uint8 HWC images + uint8 masks from pageable host memory, ToTensor + Normalize on the device
        estep(x, m)
    parallel.all_reduce_confusion(cm)
    m = Metrics(12, 11)
    m._confusion_matrix += cm.cpu().numpy()  # the batches of the end-to-end loop, all ranks
    miou = float(m.iou())
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3))
    if rank != 0:
        return
    out = {"metric": metric_name(args).replace("train", "eval"), "value": world * B * args.steps / (ms * 1e-3),
           "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic", "config": dict(workload_config(args), workload=(
               f"{args.model} eval step (forward with running statistics + argmax + confusion matrix), per-GPU batch "
               f"{B}x3x{args.height}x{args.width}")),
           "e2e": {"value": world * B * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": world * pf.h2d_bytes // args.steps,
                   "d2h_bytes_per_step": 12 * 12 * 8 / args.steps, "ms_per_step": e2e_ms / args.steps},
           "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches), "miou_random_init": miou}
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
