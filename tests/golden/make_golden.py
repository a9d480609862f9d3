"""Generates the golden fixtures in this directory by running the REFERENCE ITSELF (weiaicunzai/pytorch-camvid,
mounted read-only at /root/reference) on deterministic inputs. Run from the repository root in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own; these files pin the oracle (oracle/camvid_oracle.py) and,
through it, the CUDA path. The reference tree cannot travel to the GPU box, the fixtures can.
Shims: `np.float = float` for utils.mean_iou, sys.path += legacy/ for Metrics.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "legacy"))
np.float = float  # removed alias used at utils.py:210-213

import utils as ref_utils  # noqa: E402  (reference)
from metrics import Metrics as RefMetrics  # noqa: E402  (reference legacy/metrics.py)
from oracle import camvid_oracle as O  # noqa: E402  (only for the shared deterministic input generators)

OUT = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(0)
torch.set_num_threads(8)

KEEP_FULL = {  # small gradient tensors stored in full; every other parameter stores norm + checksum
    "unet": ["down1.0.conv.0.weight", "output.conv.0.weight", "output.conv.1.weight", "output.conv.1.bias",
             "down3.1.conv.1.weight", "up1.0.conv.1.bias", "upsample1.conv.conv.1.weight"],
    "segnet": ["encoder1.0.conv.weight", "decoder1.1.conv.weight", "decoder1.1.bn.weight", "decoder1.1.bn.bias",
               "encoder3.2.bn.weight", "decoder5.0.bn.bias"],
}


def model_fixture(name, n, h, w):
    net = ref_utils.get_model(name, 3, 12)
    sd = O.synth_state_dict(net.state_dict(), seed=1)
    net.load_state_dict(sd)
    x, t = O.synth_batch(n, h, w, seed=2)
    net.train()
    logits = net(x)
    loss = torch.nn.CrossEntropyLoss()(logits, t)
    loss.backward()
    out = {"x": x.numpy(), "target": t.numpy(), "logits": logits.detach().numpy(), "loss": loss.item()}
    names, norms, sums = [], [], []
    for k, p in net.named_parameters():
        names.append(k)
        norms.append(p.grad.double().norm().item())
        sums.append(p.grad.double().sum().item())
        if k in KEEP_FULL[name]:
            out["grad/" + k] = p.grad.numpy()
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    out["grad_sums"] = np.array(sums)
    after = net.state_dict()
    last_bn = [k for k in after if k.endswith("running_mean")][-1][:-len("running_mean")]
    first_bn = [k for k in after if k.endswith("running_mean")][0][:-len("running_mean")]
    for pre in (first_bn, last_bn):
        out["after/" + pre + "running_mean"] = after[pre + "running_mean"].numpy()
        out["after/" + pre + "running_var"] = after[pre + "running_var"].numpy()
    # eval-mode forward with the updated running statistics
    net.eval()
    with torch.no_grad():
        out["eval_logits"] = net(x).numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}_step.npz"), **out)
    print(name, "loss", out["loss"], "logits zero-fraction", float((logits == 0).float().mean()))


def metric_fixture():
    rng = np.random.default_rng(7)
    out = {}
    # case A: ordinary labels; case B: labels with out-of-range values (255 = the unused IGNORE_LABEL, and C itself)
    pred = rng.integers(0, 12, (3, 37, 41)).astype(np.int64)
    gt = rng.integers(0, 12, (3, 37, 41)).astype(np.int64)
    gt_b = gt.copy()
    gt_b[rng.random(gt.shape) < 0.05] = 255
    gt_b[rng.random(gt.shape) < 0.02] = 12
    pred_b = pred.copy()
    pred_b[rng.random(pred.shape) < 0.02] = 12
    for tag, p, g in (("a", pred, gt), ("b", pred_b, gt_b)):
        out[f"{tag}/pred"], out[f"{tag}/gt"] = p, g
        for ign in (11, 255):
            all_acc, acc, iou = ref_utils.mean_iou(torch.from_numpy(p), torch.from_numpy(g), 12, ign)
            out[f"{tag}/miou{ign}/all_acc"], out[f"{tag}/miou{ign}/acc"], out[f"{tag}/miou{ign}/iou"] = all_acc, acc, iou
            i, u, ap, al = ref_utils.intersect_and_union(p[0], g[0], 12, ign)
            out[f"{tag}/iau{ign}"] = np.stack([i, u, ap, al])
        for ign in (None, 11, 0):
            m = RefMetrics(12, ign)
            m.add(p.reshape(-1), g.reshape(-1))
            m.add(p[:1].reshape(-1), g[:1].reshape(-1))
            key = f"{tag}/metrics{ign}"
            out[key + "/cm"] = m._confusion_matrix.copy()
            out[key + "/precision"], out[key + "/recall"], out[key + "/iou"] = m.precision(), m.recall(), m.iou()
            out[key + "/iou_vec"] = m.iou(average=False)
            out[key + "/precision_vec"] = m.precision(average=False)
    # a class that never occurs -> NaN entries in acc / iou (0/0), and nan_to_num
    p0, g0 = np.clip(pred, 0, 9), np.clip(gt, 0, 9)
    all_acc, acc, iou = ref_utils.mean_iou(torch.from_numpy(p0), torch.from_numpy(g0), 12, 11)
    out["c/pred"], out["c/gt"], out["c/all_acc"], out["c/acc"], out["c/iou"] = p0, g0, all_acc, acc, iou
    _, acc2, iou2 = ref_utils.mean_iou(torch.from_numpy(p0), torch.from_numpy(g0), 12, 11, nan_to_num=-1)
    out["c/acc_n2n"], out["c/iou_n2n"] = acc2, iou2
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics fixtures written")


def loss_fixture():
    rng = np.random.default_rng(9)
    logits = torch.from_numpy(np.maximum(rng.standard_normal((2, 12, 9, 13)) * 2, 0).astype(np.float32))
    target = torch.from_numpy(rng.integers(0, 12, (2, 9, 13)).astype(np.int64))
    out = {"logits": logits.numpy(), "target": target.numpy()}
    for ign in (-100, 11):
        lg = logits.clone().requires_grad_(True)
        loss = torch.nn.CrossEntropyLoss(ignore_index=ign)(lg, target)
        loss.backward()
        out[f"loss{ign}"], out[f"grad{ign}"] = loss.item(), lg.grad.numpy()
    out["argmax"] = logits.argmax(dim=1).numpy()
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)
    print("loss fixtures written")


def input_stage_fixture():
    """transforms.ToTensor + transforms.Normalize (transforms.py:485-538) with the CamVid statistics
    (conf/settings.py:8-9) on uint8 HWC images that contain every byte value in every channel, plus the mask's
    .long(): the reference output the device input stage (cvb_input_stage_u8) must reproduce bit for bit."""
    import transforms as ref_transforms  # reference
    from conf import settings  # reference
    rng = np.random.default_rng(11)
    out = {"mean": np.array(settings.MEAN), "std": np.array(settings.STD)}
    for tag, (n, h, w) in (("vec", (2, 24, 32)), ("odd", (3, 9, 11))):  # h*w % 4 == 0 / generic kernel
        img = rng.integers(0, 256, (n, h, w, 3)).astype(np.uint8)
        img.reshape(-1, 3)[:256] = np.stack([np.arange(256), np.arange(255, -1, -1), np.roll(np.arange(256), 97)], 1)
        mask = rng.integers(0, 12, (n, h, w)).astype(np.uint8)
        chain = ref_transforms.Compose([ref_transforms.ToTensor(), ref_transforms.Normalize(settings.MEAN, settings.STD)])
        res = [chain(img[i].copy(), mask[i].copy()) for i in range(n)]
        out[f"{tag}/img"], out[f"{tag}/mask"] = img, mask
        out[f"{tag}/out"] = torch.stack([r[0] for r in res]).numpy()
        out[f"{tag}/mask_out"] = torch.stack([r[1] for r in res]).numpy()
        assert out[f"{tag}/out"].dtype == np.float32 and out[f"{tag}/mask_out"].dtype == np.int64
    np.savez_compressed(os.path.join(OUT, "input_stage.npz"), **out)
    print("input stage fixtures written")


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "models" in only:
        model_fixture("unet", 2, 40, 72)
        model_fixture("segnet", 2, 40, 72)
    if not only or "metrics" in only:
        metric_fixture()
    if not only or "loss" in only:
        loss_fixture()
    if not only or "input_stage" in only:
        input_stage_fixture()
