"""CPU tests: the oracle restatement (oracle/camvid_oracle.py) against the golden fixtures generated from the
reference itself (tests/golden/make_golden.py), and against the live reference when /root/reference is mounted."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import camvid_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"


def _template(name):
    import camvid_b200  # noqa: F401
    from camvid_b200.utils import get_model
    return get_model(name, 3, 12).state_dict()


@pytest.mark.parametrize("name", ["unet", "segnet"])
def test_model_step_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLD, f"{name}_step.npz"))
    sd = O.synth_state_dict(_template(name), seed=1)
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    xs, ts = O.synth_batch(*x.shape[:1], *x.shape[2:], seed=2)
    assert torch.equal(xs, x) and torch.equal(ts, t)  # the generators are deterministic across machines
    loss, logits, grads, after = O.train_step(name, sd, x, t)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-4)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    names = list(g["grad_names"])
    for k, nrm in zip(names, g["grad_norms"]):
        tol = 1e-5 + 2e-3 * nrm
        assert abs(grads[k].double().norm().item() - nrm) < tol, k
    for key in g.files:
        if key.startswith("grad/"):
            np.testing.assert_allclose(grads[key[5:]].numpy(), g[key], rtol=2e-3, atol=2e-5)
        if key.startswith("after/"):
            np.testing.assert_allclose(after[key[6:]].numpy(), g[key], rtol=1e-4, atol=1e-6)
    ev = O.FORWARD[name](after, x, train=False)
    np.testing.assert_allclose(ev.numpy(), g["eval_logits"], rtol=1e-4, atol=1e-4)


def test_metrics_match_reference_fixture():
    g = np.load(os.path.join(GOLD, "metrics.npz"))
    for tag in ("a", "b"):
        p, t = g[f"{tag}/pred"], g[f"{tag}/gt"]
        for ign in (11, 255):
            all_acc, acc, iou = O.mean_iou(p, t, 12, ign)
            assert all_acc == g[f"{tag}/miou{ign}/all_acc"]
            np.testing.assert_array_equal(acc, g[f"{tag}/miou{ign}/acc"])
            np.testing.assert_array_equal(iou, g[f"{tag}/miou{ign}/iou"])
            np.testing.assert_array_equal(np.stack(O.intersect_and_union(p[0], t[0], 12, ign)), g[f"{tag}/iau{ign}"])
        for ign in (None, 11, 0):
            m = O.Metrics(12, ign)
            m.add(p.reshape(-1), t.reshape(-1))
            m.add(p[:1].reshape(-1), t[:1].reshape(-1))
            key = f"{tag}/metrics{ign}"
            np.testing.assert_array_equal(m.cm, g[key + "/cm"])
            assert m.precision() == g[key + "/precision"] and m.recall() == g[key + "/recall"]
            assert m.iou() == g[key + "/iou"]
            np.testing.assert_array_equal(m.iou(average=False), g[key + "/iou_vec"])
            np.testing.assert_array_equal(m.precision(average=False), g[key + "/precision_vec"])
    all_acc, acc, iou = O.mean_iou(g["c/pred"], g["c/gt"], 12, 11)
    assert all_acc == g["c/all_acc"]
    np.testing.assert_array_equal(acc, g["c/acc"])  # NaN == NaN under assert_array_equal
    np.testing.assert_array_equal(iou, g["c/iou"])
    _, acc2, iou2 = O.mean_iou(g["c/pred"], g["c/gt"], 12, 11, nan_to_num=-1)
    np.testing.assert_array_equal(acc2, g["c/acc_n2n"])
    np.testing.assert_array_equal(iou2, g["c/iou_n2n"])


def test_loss_matches_reference_fixture():
    g = np.load(os.path.join(GOLD, "loss.npz"))
    for ign in (-100, 11):
        lg = torch.from_numpy(g["logits"]).requires_grad_(True)
        loss = O.cross_entropy(lg, torch.from_numpy(g["target"]), ign)
        loss.backward()
        assert abs(loss.item() - float(g[f"loss{ign}"])) < 1e-6
        np.testing.assert_allclose(lg.grad.numpy(), g[f"grad{ign}"], rtol=1e-5, atol=1e-8)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", ["unet", "segnet"])
def test_oracle_matches_live_reference(name):
    """Random torch-default init, odd sizes (pad / output_size paths), live reference forward+backward."""
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import importlib
    ref_utils = importlib.import_module("utils")
    try:
        torch.manual_seed(3)
        net = ref_utils.get_model(name, 3, 12)
        x, t = O.synth_batch(1, 45, 61 if name == "unet" else 70, seed=4)
        net.train()
        logits = net(x)
        loss = torch.nn.CrossEntropyLoss()(logits, t)
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        # undo the running-stat update the forward just made so the oracle starts from the same state
        torch.manual_seed(3)
        sd0 = {k: v.clone() for k, v in ref_utils.get_model(name, 3, 12).state_dict().items()}
        loss.backward()
        o_loss, o_logits, o_grads, o_after = O.train_step(name, sd0, x, t)
        assert torch.allclose(o_logits, logits.detach(), rtol=1e-4, atol=1e-5)
        assert abs(o_loss.item() - loss.item()) < 1e-5
        for k, p in net.named_parameters():
            assert torch.allclose(o_grads[k], p.grad, rtol=1e-3, atol=1e-5), k
        for k, v in net.state_dict().items():
            assert torch.allclose(o_after[k].float(), v.float(), rtol=1e-4, atol=1e-6), k
    finally:
        sys.path.remove(REF)
        for mod in ("utils", "models", "models.unet", "models.segnet"):
            sys.modules.pop(mod, None)
