"""ORACLE -- test infrastructure, not product code.

CPU/fp32 restatement of the reference hot path (weiaicunzai/pytorch-camvid): the UNet and SegNet forward passes,
the loss, and the two metric conventions. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product (camvid_b200) never does.

Where the arithmetic lives: the reference is pure Python over a third-party dependency, PyTorch (unpinned in the
reference; 2.11.0+cu128 here) plus numpy / scikit-learn for the metrics. This file restates the reference's call
sequence as plain functional torch fp32 ops and numpy so it can travel to the GPU box (the reference tree cannot).

Pinning: the reference ships no tests, fixtures or golden vectors, so the oracle is pinned
against the reference ITSELF: tests/golden/make_golden.py imports /root/reference, runs its modules on seeded inputs
and commits the outputs as fixtures; tests/test_oracle.py checks this restatement against those fixtures everywhere
and against the live reference whenever /root/reference is present.

State-dict keys follow the reference exactly (UNet: `down1.0.conv.0.weight`, SegNet: `encoder1.0.conv.weight`).
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# building block: conv3x3(pad 1, bias) -> BatchNorm2d -> ReLU   (models/unet.py:5-17, models/segnet.py:5-17)
# ----------------------------------------------------------------------------------------------------------------
RECORD = None  # tests may set this to a dict; filled with block name -> {"x", "out"} and, by train_step, {"dx", "dout"}
STORAGE = "fp32"  # "fp32": the reference's arithmetic. "bf16": the same network with the north_star's storage format


def _cbr(x, sd, conv, bn, train, momentum=0.1, eps=1e-5):
    impl = _cbr_impl if STORAGE == "fp32" else _cbr_bf16
    out = impl(x, sd, conv, bn, train, momentum, eps)
    if RECORD is not None:
        RECORD[conv.rsplit(".conv", 1)[0] if conv.endswith(".conv") else conv[:-len(".conv.0")]] = {"x": x, "out": out}
    return out


def _cbr_impl(x, sd, conv, bn, train, momentum=0.1, eps=1e-5):
    y = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], padding=1)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     train, momentum, eps)
    if train and (bn + ".num_batches_tracked") in sd:
        sd[bn + ".num_batches_tracked"] += 1
    return F.relu(y)


# ----------------------------------------------------------------------------------------------------------------
# bf16 storage model (STORAGE = "bf16").  BASELINE.json's north_star fixes the format of the accelerated path:
# "NHWC bf16 with fp32 accumulate".  This is the reference network with exactly that and nothing else changed:
# conv operands (activations, weights), conv outputs, activations and activation gradients are HELD in bf16 (rounded
# to nearest-even at the point they would be stored); every sum, statistic, normalisation, interpolation and the loss
# is fp32, like the reference.  It exists because the network is chaotic at random init (a perturbation grows about
# x1.2 per conv+BN+ReLU block, tests/precision_sim.py), so ANY bf16 implementation -- including the reference under
# torch autocast -- sits ~1e-1 away from the fp32 logits; the CUDA path is checked tightly against this model and,
# block by block on identical inputs, against the fp32 reference.
# ----------------------------------------------------------------------------------------------------------------
def _r(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _StoreBF16(torch.autograd.Function):
    """y = bf16(x) in forward and/or g = bf16(g) in backward (a tensor and its gradient written to a bf16 buffer)."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return _r(x) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (_r(g) if ctx.bwd else g), None, None


class _BlockBF16(torch.autograd.Function):
    """conv3x3 -> BatchNorm(batch statistics) -> ReLU with bf16 storage; backward restated with the same storage
    points (dy and dx held in bf16, reductions and parameter gradients fp32)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, eps):
        x16, w16 = _r(x), _r(w)
        y32 = F.conv2d(x16, w16, None, padding=1)  # the bias cancels under batch statistics
        n = y32.numel() // y32.shape[1]
        mean = y32.double().mean((0, 2, 3))
        var = (y32.double() ** 2).mean((0, 2, 3)) - mean ** 2
        invstd = (1.0 / torch.sqrt(var.clamp_min(0) + eps)).float()
        mean = mean.float()
        scale = gamma * invstd
        shift = beta - mean * scale
        y16 = _r(y32)
        z = y16 * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        ctx.save_for_backward(x16, w16, y16, z > 0, mean, invstd, gamma)
        ctx.n = n
        ctx.mark_non_differentiable(mean, var)
        return _r(F.relu(z)), mean, var.float()

    @staticmethod
    def backward(ctx, da, _m, _v):
        x16, w16, y16, pos, mean, invstd, gamma = ctx.saved_tensors
        n = ctx.n
        g = da * pos
        sg = g.double().sum((0, 2, 3))
        sgy = (g.double() * y16.double()).sum((0, 2, 3))
        dgamma = invstd.double() * (sgy - mean.double() * sg)
        dbeta = sg
        sc = (gamma * invstd).double()
        c0 = sc.float().view(1, -1, 1, 1)
        c1 = (-sc * invstd.double() * dgamma / n).float().view(1, -1, 1, 1)
        c2 = (-sc * dbeta / n + sc * mean.double() * invstd.double() * dgamma / n).float().view(1, -1, 1, 1)
        dy16 = _r(g * c0 + (y16 * c1 + c2))
        dw = torch.nn.grad.conv2d_weight(x16, w16.shape, dy16, padding=1)
        dx = _r(torch.nn.grad.conv2d_input(x16.shape, w16, dy16, padding=1)) if ctx.needs_input_grad[0] else None
        return dx, dw, dgamma.float(), dbeta.float(), None


def _cbr_bf16(x, sd, conv, bn, train, momentum=0.1, eps=1e-5):
    w, b = sd[conv + ".weight"], sd[conv + ".bias"]
    gamma, beta = sd[bn + ".weight"], sd[bn + ".bias"]
    if not train:  # BatchNorm folded into the conv epilogue: one rounding, of the activation
        s = gamma * torch.rsqrt(sd[bn + ".running_var"] + eps)
        acc = F.conv2d(_r(x), _r(w), None, padding=1)
        return _r(F.relu(acc * s.view(1, -1, 1, 1) + (beta + (b - sd[bn + ".running_mean"]) * s).view(1, -1, 1, 1)))
    a, mean, var = _BlockBF16.apply(x, w, gamma, beta, eps)
    with torch.no_grad():
        n = a.numel() // a.shape[1]
        sd[bn + ".running_mean"].mul_(1 - momentum).add_(momentum * (mean + b.detach()))
        sd[bn + ".running_var"].mul_(1 - momentum).add_(momentum * var * (n / max(n - 1, 1)))
        if (bn + ".num_batches_tracked") in sd:
            sd[bn + ".num_batches_tracked"] += 1
    return _StoreBF16.apply(a, False, True)  # gradients from several consumers are summed, then stored in bf16


def block_step(x, w, bias, gamma, beta, dout, need_dx=True, storage="fp32", eps=1e-5):
    """One conv3x3(pad 1, bias) -> BatchNorm2d(batch statistics) -> ReLU block (models/unet.py:5-17,
    models/segnet.py:5-17) forward + backward on its own: returns (a, dx or None, dw, dgamma, dbeta) for the output
    gradient `dout`. storage="bf16": the bf16 storage model of the block (see above). Used by the teacher-forced block
    tests at the BASELINE geometries, where running the whole network through autograd on the CPU is too large."""
    xr = x.detach().clone().requires_grad_(need_dx)
    ps = [t.detach().clone().requires_grad_(True) for t in (w, gamma, beta)]
    if storage == "fp32":
        y = F.conv2d(xr, ps[0], bias, padding=1)
        a = F.relu(F.batch_norm(y, None, None, ps[1], ps[2], True, 0.1, eps))
        gout = dout
    else:
        a, _, _ = _BlockBF16.apply(xr, ps[0], ps[1], ps[2], eps)
        gout = _r(dout)
    g = torch.autograd.grad(a, ps + ([xr] if need_dx else []), grad_outputs=gout)
    return a.detach(), (g[3] if need_dx else None), g[0], g[1], g[2]


def _held(x, fwd=True, bwd=True):
    """Marks a tensor (and its gradient) as written to a bf16 buffer; identity under fp32 storage."""
    return x if STORAGE == "fp32" else _StoreBF16.apply(x, fwd, bwd)


def _unet_block(x, sd, prefix, train):
    """BasicConv2d: Sequential(conv, bn, relu) stored under `<prefix>.conv.{0,1}` (models/unet.py:10-14)."""
    return _cbr(x, sd, prefix + ".conv.0", prefix + ".conv.1", train)


def unet_forward(sd, x, train=True):
    """models/unet.py:94-156. `sd`: dict name -> tensor (running stats are updated in place when train)."""
    skips = []
    x = _held(x, True, False)
    for level in range(1, 6):
        x = _unet_block(x, sd, f"down{level}.0", train)
        x = _unet_block(x, sd, f"down{level}.1", train)
        if level < 5:
            skips.append(x)
            x = F.max_pool2d(x, 2, 2)  # models/unet.py:92,100-109
    for stage in range(1, 5):
        skip = skips[4 - stage]
        x = _held(F.interpolate(_held(x, False, True), scale_factor=2, mode="bilinear", align_corners=True))  # :25,29
        x = _unet_block(x, sd, f"upsample{stage}.conv", train)
        dh, dw = skip.size(2) - x.size(2), skip.size(3) - x.size(3)
        x = F.pad(x, [dw // 2, dw - dw // 2, dh // 2, dh - dh // 2])  # models/unet.py:120-123
        x = torch.cat([x, skip], dim=1)  # upsampled branch first, models/unet.py:124
        x = _unet_block(x, sd, f"up{stage}.0", train)
        x = _unet_block(x, sd, f"up{stage}.1", train)
    return _held(_unet_block(x, sd, "output", train), False, True)  # models/unet.py:91,154 (logits are post-ReLU)


SEGNET_DEPTH = (2, 2, 3, 3, 3)


def segnet_forward(sd, x, train=True, return_indices=False):
    """models/segnet.py:82-119."""
    shapes, indices = [], []
    x = _held(x, True, False)
    for s in range(1, 6):
        for j in range(SEGNET_DEPTH[s - 1]):
            x = _cbr(x, sd, f"encoder{s}.{j}.conv", f"encoder{s}.{j}.bn", train)
        shapes.append(x.shape)
        x, idx = F.max_pool2d(x, 2, return_indices=True)  # models/segnet.py:79
        indices.append(idx)
    for s in range(5, 0, -1):
        x = F.max_unpool2d(x, indices[s - 1], 2, output_size=shapes[s - 1])  # models/segnet.py:80,104-116
        for j in range(SEGNET_DEPTH[s - 1]):
            x = _cbr(x, sd, f"decoder{s}.{j}.conv", f"decoder{s}.{j}.bn", train)
    x = _held(x, False, True)
    return (x, indices) if return_indices else x


FORWARD = {"unet": unet_forward, "segnet": segnet_forward}


def cross_entropy(logits, target, ignore_index=-100):
    """nn.CrossEntropyLoss() as constructed at train.py:105 / eval.py:42 (mean over non-ignored pixels)."""
    return F.cross_entropy(logits, target, ignore_index=ignore_index)


def train_step(name, sd, x, target, ignore_index=-100, storage="fp32"):
    """One forward + backward of train.py:128-131 on a copy of `sd`. Returns (loss, logits, grads, new_sd).
    storage="bf16" runs the bf16 storage model of the same network (see above)."""
    global STORAGE
    prev, STORAGE = STORAGE, storage
    try:
        return _train_step(name, sd, x, target, ignore_index)
    finally:
        STORAGE = prev


def forward(name, sd, x, train=False, storage="fp32"):
    global STORAGE
    prev, STORAGE = STORAGE, storage
    try:
        with torch.no_grad():
            return FORWARD[name](sd, x, train)
    finally:
        STORAGE = prev


def _train_step(name, sd, x, target, ignore_index):
    sd = {k: v.clone() for k, v in sd.items()}
    leaves = {}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            leaves[k] = sd[k] = v.detach().requires_grad_(True)
    logits = FORWARD[name](sd, x, True)
    loss = cross_entropy(logits, target, ignore_index)
    extra = []
    if RECORD is not None:  # also differentiate w.r.t. every block's input and output (teacher-forced block tests)
        for rec in RECORD.values():
            extra += [rec["out"]] + ([rec["x"]] if rec["x"].requires_grad else [])
    grads = torch.autograd.grad(loss, list(leaves.values()) + extra, allow_unused=True)
    if RECORD is not None:
        it = iter(grads[len(leaves):])
        for rec in RECORD.values():
            rec["dout"] = next(it)
            rec["dx"] = next(it) if rec["x"].requires_grad else None
            rec["x"], rec["out"] = rec["x"].detach(), rec["out"].detach()
    grads = grads[:len(leaves)]
    grads = [torch.zeros_like(p) if g is None else g for p, g in zip(leaves.values(), grads)]  # bf16 model: conv bias
    return loss.detach(), logits.detach(), dict(zip(leaves.keys(), grads)), {k: v.detach() for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------------------------
# input stage: transforms.ToTensor + transforms.Normalize (transforms.py:485-538)
# ----------------------------------------------------------------------------------------------------------------
def to_tensor_normalize(img_u8_hwc, mask_u8, mean, std):
    """img uint8 [N,H,W,C] (HWC, cv2) -> fp32 [N,C,H,W]; mask uint8 [N,H,W] -> int64. numpy fp32, operation by
    operation as the reference: `img.float() / 255.0` (transforms.py:500-501: IEEE division), then
    `img.sub_(mean[:, None, None]).div_(std[:, None, None])` (transforms.py:536-538)."""
    x = np.asarray(img_u8_hwc).transpose(0, 3, 1, 2).astype(np.float32)
    x = x / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(1, -1, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, -1, 1, 1)
    x = (x - m).astype(np.float32) / s
    return x.astype(np.float32), np.asarray(mask_u8).astype(np.int64)


# ----------------------------------------------------------------------------------------------------------------
# metrics
# ----------------------------------------------------------------------------------------------------------------
def _hist(values, num_classes):
    """np.histogram(values, bins=np.arange(C+1)) restated: unit bins, the last one closed on the right."""
    v = np.asarray(values).reshape(-1).astype(np.int64)
    out = np.bincount(v[(v >= 0) & (v < num_classes)], minlength=num_classes)[:num_classes].astype(np.int64)
    out[num_classes - 1] += int((v == num_classes).sum())
    return out


def intersect_and_union(pred_label, label, num_classes, ignore_index):
    """utils.py:162-190."""
    pred_label, label = np.asarray(pred_label), np.asarray(label)
    keep = label != ignore_index
    pred_label, label = pred_label[keep], label[keep]
    area_intersect = _hist(pred_label[pred_label == label], num_classes)
    area_pred = _hist(pred_label, num_classes)
    area_label = _hist(label, num_classes)
    return area_intersect, area_pred + area_label - area_intersect, area_pred, area_label


def mean_iou(results, gt_seg_maps, num_classes, ignore_index, nan_to_num=None):
    """utils.py:193-228 (with the removed `np.float` alias read as float64)."""
    assert len(results) == len(gt_seg_maps)
    tot = np.zeros((4, num_classes), dtype=np.float64)
    for r, g in zip(results, gt_seg_maps):
        tot += np.stack(intersect_and_union(r, g, num_classes, ignore_index)).astype(np.float64)
    inter, union, _, lab = tot
    with np.errstate(divide="ignore", invalid="ignore"):
        all_acc, acc, iou = inter.sum() / lab.sum(), inter / lab, inter / union
    if nan_to_num is not None:
        acc, iou = np.nan_to_num(acc, nan=nan_to_num), np.nan_to_num(iou, nan=nan_to_num)
    return all_acc, acc, iou


def confusion_matrix(gts, preds, num_classes):
    """sklearn.metrics.confusion_matrix(gts, preds, labels=range(C)) as used at legacy/metrics.py:29:
    rows = ground truth, cols = prediction, pairs with a label outside range(C) dropped."""
    g = np.asarray(gts).reshape(-1).astype(np.int64)
    p = np.asarray(preds).reshape(-1).astype(np.int64)
    ok = (g >= 0) & (g < num_classes) & (p >= 0) & (p < num_classes)
    return np.bincount(g[ok] * num_classes + p[ok], minlength=num_classes ** 2).reshape(num_classes, num_classes)


class Metrics:
    """legacy/metrics.py:6-71."""

    def __init__(self, class_num, ignore_index=None):
        self.class_num, self.ignore_index = class_num, ignore_index
        self.cm = np.zeros((class_num, class_num))

    def add(self, preds, gts):
        self.cm += confusion_matrix(gts, preds, self.class_num)

    def clear(self):
        self.cm.fill(0)

    def _select(self, v, drop, average):
        if drop:
            v = v[[i for i in range(self.class_num) if i != self.ignore_index]]
        return v.mean() if average else v

    def precision(self, average=True):
        return self._select(np.diag(self.cm) / (self.cm.sum(0) + 1e-15), bool(self.ignore_index), average)

    def recall(self, average=True):
        return self._select(np.diag(self.cm) / (self.cm.sum(1) + 1e-15), bool(self.ignore_index), average)

    def iou(self, average=True):
        d = np.diag(self.cm)
        return self._select(d / (self.cm.sum(1) + self.cm.sum(0) - d + 1e-15), True, average)


# ----------------------------------------------------------------------------------------------------------------
# deterministic parameters / inputs shared by the golden generator and the tests (no torch RNG involved)
# ----------------------------------------------------------------------------------------------------------------
def synth_state_dict(template_sd, seed=0):
    """Fills a state dict (same keys / shapes as the reference module's) from a name-keyed numpy generator:
    conv weights ~ U(+-1/sqrt(fan_in)), conv bias small, BN gamma in [0.5, 1.5], beta small, running stats
    perturbed -- non-trivial everywhere so every term of the forward/backward is exercised."""
    import zlib
    out = {}
    for k, v in template_sd.items():
        rng = np.random.default_rng([seed, zlib.crc32(k.encode())])
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.int64)
        elif v.dim() == 4:
            b = 1.0 / np.sqrt(shape[1] * 9)
            out[k] = torch.from_numpy(rng.uniform(-b, b, shape).astype(np.float32))
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy(rng.normal(0, 0.1, shape).astype(np.float32))
        elif k.endswith("bias"):
            out[k] = torch.from_numpy(rng.normal(0, 0.1, shape).astype(np.float32))
        else:  # BN weight
            out[k] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
    return out


def synth_batch(n, h, w, classes=12, seed=0):
    rng = np.random.default_rng([seed, n, h, w])
    x = torch.from_numpy(rng.standard_normal((n, 3, h, w)).astype(np.float32))
    t = torch.from_numpy(rng.integers(0, classes, (n, h, w)).astype(np.int64))
    return x, t
