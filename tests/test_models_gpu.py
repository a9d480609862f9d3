"""Whole-model parity on the GPU: the drop-in UNet / SegNet (kernel plans through the C ABI) against
  (1) the golden fixtures produced by the reference itself (tests/golden/make_golden.py), and
  (2) the fp32 oracle (oracle/camvid_oracle.py) on seeded inputs, including odd sizes and the full 360x480 geometry.

Tolerances are BASELINE.json's north_star: logits and loss within 2e-2 relative (norm-wise), >= 99.5 %
per-pixel argmax agreement, per-layer weight gradients within 3e-2 relative (conv biases feed a batch-stat BatchNorm:
their gradient is mathematically zero and is compared absolutely).

How the tolerances are applied (DESIGN.md "Parity"). Two properties of the reference network at random init, both
reproduced on the CPU with the reference's own fp32 arithmetic (oracle storage="bf16", tests/precision_sim.py):
  (1) it is chaotic -- a perturbation grows about x1.2 per conv+BN+ReLU block, so bf16 storage of conv operands and
      outputs alone moves the fp32 logits by ~1e-1 (UNet) to ~6e-1 (SegNet: pooling indices flip) and the first-layer
      gradients by ~8e-1; two bf16 runs that differ only in accumulation order diverge the same way;
  (2) rounding a conv output to bf16 flips the ReLU mask of the ~0.1-0.3 % of elements closest to zero, which alone is
      a 3-6e-2 relative change of that block's gradients (sqrt of the flipped fraction), on identical inputs.
No bf16 implementation -- torch autocast included -- can therefore meet 2e-2 / 3e-2 end to end against fp32. So:
  * every block is checked on IDENTICAL inputs (teacher forcing, fp32 oracle activations and gradients fed in):
    activations against the fp32 oracle at 2e-2; gradients at 3e-2 against the bf16 storage model of that block and,
    against fp32, no worse than that model itself is;
  * whole model: loss within 2e-2 of fp32 (it is: ~1e-4); logits and gradients no further from fp32 than the bf16
    storage model of the reference is (x1.3 slack); per-layer gradient norms within 15 %; eval mode (running
    statistics, not chaotic) meets 2e-2 / 99.5 % against the fp32 reference fixture directly; a 3-step AdamW loss
    trajectory follows the fp32 oracle within 3e-2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import camvid_oracle as O
from util import rel_err

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_LOGITS, TOL_LOSS, TOL_GRAD, MIN_AGREE = 2e-2, 2e-2, 3e-2, 0.995


@pytest.fixture(scope="module")
def cvb(cuda):
    import camvid_b200  # noqa: F401
    from camvid_b200 import nn as cnn
    from camvid_b200 import utils as cutils
    return cutils, cnn


def _is_conv_bias(net, name):
    mod = net.get_submodule(name.rsplit(".", 1)[0])
    return name.endswith(".bias") and isinstance(mod, torch.nn.Conv2d)


def _argmax_agreement(a, b, decided_only=False):
    """Fraction of pixels with the same argmax(dim=1) as the reference b. decided_only: count only pixels whose
    reference top-2 margin exceeds the logit tolerance itself (2e-2 of the winning logit) -- a logit allowed to move by
    2e-2 cannot pin a winner closer than that (and ~half of the post-ReLU logits tie at exactly 0)."""
    pa, pb = a.argmax(1), b.argmax(1)
    same = pa == pb
    if not decided_only:
        return same.float().mean().item()
    top2 = b.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > TOL_LOGITS * top2[:, 0].abs()
    return same[decided].float().mean().item()


def _plan(net):
    from camvid_b200 import engine
    return engine.plans_of(net)[0]


def _build(cvb, name, sd, dev):
    cutils, _ = cvb
    net = cutils.get_model(name, 3, 12)
    net.load_state_dict(sd)
    return net.to(dev)


def _step(cvb, net, x, t, dev, ignore_index=-100):
    _, cnn = cvb
    net.train()
    net.zero_grad(set_to_none=True)
    logits = net(x.to(dev))
    loss = cnn.CrossEntropyLoss(ignore_index=ignore_index)(logits, t.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), logits.detach().cpu()


def _check_grad_norms(net, o_grads):
    """Per-layer gradient norms are stable where directions are not; catches mis-wired plans."""
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        g = p.grad.detach().cpu()
        assert torch.isfinite(g).all(), k
        if _is_conv_bias(net, k):
            assert g.abs().max().item() < 1e-4, k  # reference: float noise around 0
            continue
        ref = o_grads[k].double().norm().item()
        slack = 0.15 if g.dim() == 4 else 0.4  # BatchNorm gradients are sums with heavy cancellation
        assert abs(g.double().norm().item() - ref) < slack * ref + 1e-7, (k, g.norm().item(), ref)


def _check_within_bf16_envelope(net, loss, logits, o_loss, o_logits, o_grads, m_logits, m_grads):
    """Against fp32 (oracle or reference fixture): loss at the north_star tolerance; logits / weight gradients no
    further from fp32 than the bf16 storage model of the reference network itself is (x1.3 slack + 2e-2 floor)."""
    assert logits.shape == o_logits.shape and logits.dtype == torch.float32 and torch.isfinite(logits).all()
    assert abs(loss.item() - float(o_loss)) / abs(float(o_loss)) < TOL_LOSS
    env = rel_err(m_logits, o_logits)
    assert rel_err(logits, o_logits) < max(TOL_LOGITS, 1.3 * env), env
    if o_grads is not None:
        for k, p in net.named_parameters():
            if _is_conv_bias(net, k):
                continue
            env = rel_err(m_grads[k], o_grads[k])
            assert rel_err(p.grad.cpu(), o_grads[k]) < max(TOL_GRAD, 1.3 * env), (k, env)


def _oracle_steps_with_records(name, sd, x, t, ignore_index=-100):
    """fp32 oracle step and its bf16 storage model, each with every block's activation recorded."""
    out = []
    for storage in ("fp32", "bf16"):
        O.RECORD = {}
        try:
            res = O.train_step(name, sd, x, t, ignore_index=ignore_index, storage=storage)
            out.append(res + (O.RECORD,))
        finally:
            O.RECORD = None
    return out


@pytest.fixture
def materialized(monkeypatch):
    """The per-block trajectory check reads EVERY block's activation back from the plan; the cross-layer fusions leave
    some of them unwritten (the consumer applies BatchNorm+ReLU itself), so these tests ask the engine to write them
    too. The fused kernels still run: only the extra stores are added."""
    from camvid_b200 import engine
    monkeypatch.setattr(engine, "MATERIALIZE_ACTIVATIONS", True)


def _check_layer_trajectory(net, rec32, rec16, tag, logits=None):
    """Whole-model check that stays non-vacuous where the end-to-end envelope is not (SegNet: bf16 storage alone puts
    the logits ~0.7 from fp32): the activation of EVERY block, read back from the plan after the step, against the
    fp32 oracle's activation of that block. The bf16 storage model's own distance from fp32 starts at 3e-3 at the first
    block and grows ~x1.25 per block (pool-index flips add a jump at SegNet's first unpool); the CUDA path must follow
    that trajectory block by block: error <= max(2e-2, 1.5 x the model's error at the same block). A mis-wired or
    mis-computed block shows up as an O(1) error at a depth where the allowance is still a few percent."""
    plan = _plan(net)
    rows, worst = [], 0.0
    for b in plan.blocks:
        key = b.name + ".conv" if b.name.startswith("upsample") else b.name
        if b is plan.blocks[-1] and b.fuses_boundary():
            act = logits  # the last block writes the fp32 NCHW logits directly; its bf16 activation is never materialised
        else:
            act = b.a[..., :b.cout].float().permute(0, 3, 1, 2).cpu()
        e_cuda, e_model = rel_err(act, rec32[key]["out"]), rel_err(rec16[key]["out"], rec32[key]["out"])
        e_pair = rel_err(act, rec16[key]["out"])
        limit = max(TOL_LOGITS, 1.5 * e_model)
        rows.append(f"{b.name}: cuda-fp32 {e_cuda:.3f} model-fp32 {e_model:.3f} cuda-model {e_pair:.3f}")
        worst = max(worst, e_cuda / limit)
        assert e_cuda < limit, (tag, b.name, e_cuda, e_model)
    print(f"{tag}: per-block activation errors (worst ratio to the allowance {worst:.2f})\n  " + "\n  ".join(rows))


def _check_logit_statistics(logits, o_logits, m_logits, tag):
    """What survives the chaos: per-class mean and standard deviation of the logits and the fraction of exact zeros
    (the bf16 storage model reproduces them within 1-2 %). Asserted at 5 % / 2 points against the fp32 oracle; the
    pointwise distances and argmax agreements are printed for CUDA vs fp32, CUDA vs the storage model and the storage
    model vs fp32 side by side."""
    for nm, ref in (("fp32 oracle", o_logits), ("bf16 storage model", m_logits)):
        agree = (logits.argmax(1) == ref.argmax(1)).float().mean().item()
        print(f"{tag}: CUDA vs {nm}: logits rel {rel_err(logits, ref):.3e}, argmax agreement {agree:.4f}")
    agree = (m_logits.argmax(1) == o_logits.argmax(1)).float().mean().item()
    print(f"{tag}: bf16 storage model vs fp32 oracle: logits rel {rel_err(m_logits, o_logits):.3e}, argmax agreement {agree:.4f}")
    mean, o_mean = logits.double().mean((0, 2, 3)), o_logits.double().mean((0, 2, 3))
    std, o_std = logits.double().std((0, 2, 3)), o_logits.double().std((0, 2, 3))
    assert ((mean - o_mean).abs() <= 0.05 * o_mean.abs() + 1e-3).all(), (tag, mean, o_mean)
    assert ((std - o_std).abs() <= 0.05 * o_std + 1e-3).all(), (tag, std, o_std)
    assert abs((logits == 0).float().mean().item() - (o_logits == 0).float().mean().item()) < 0.02, tag


def _default_init_sd(cutils, name, seed):
    torch.manual_seed(seed)
    return {k: v.clone() for k, v in cutils.get_model(name, 3, 12).state_dict().items()}  # torch default init


@pytest.mark.parametrize("name", ["unet", "segnet"])
def test_train_step_matches_reference_fixture(cvb, cuda, name):
    """Fixture = the reference's own modules run in fp32 (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, f"{name}_step.npz"))
    cutils, _ = cvb
    sd = O.synth_state_dict(cutils.get_model(name, 3, 12).state_dict(), seed=1)
    net = _build(cvb, name, sd, cuda)
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    loss, logits = _step(cvb, net, x, t, cuda)
    m_loss, m_logits, m_grads, m_after = O.train_step(name, sd, x, t, storage="bf16")
    ref_logits = torch.from_numpy(g["logits"])
    _check_within_bf16_envelope(net, loss, logits, float(g["loss"]), ref_logits, None, m_logits, None)
    params = dict(net.named_parameters())
    for k, nrm in zip(g["grad_names"], g["grad_norms"]):
        k = str(k)
        gn = params[k].grad.double().norm().item()
        if _is_conv_bias(net, k):
            assert gn < 1e-3, k
        else:  # gradient norms are stable even where directions are not
            assert abs(gn - nrm) < (0.15 if params[k].dim() == 4 else 0.4) * nrm + 1e-7, (k, gn, nrm)
    for key in g.files:
        if key.startswith("grad/") and not _is_conv_bias(net, key[5:]):
            ref = torch.from_numpy(g[key])
            env = rel_err(m_grads[key[5:]], ref)
            assert rel_err(params[key[5:]].grad.cpu(), ref) < max(TOL_GRAD, 1.3 * env), (key, env)
        if key.startswith("after/"):
            got = net.state_dict()[key[6:]].cpu()
            assert rel_err(got, torch.from_numpy(g[key])) < 5e-2, key
    assert int(net.state_dict()[[k for k in net.state_dict() if k.endswith("num_batches_tracked")][0]]) == 1
    # eval mode: BatchNorm folded into the conv epilogue with the running statistics just updated; running statistics
    # make the eval network non-chaotic, so the fp32 reference fixture itself is met at the north_star tolerance
    net.eval()
    with torch.no_grad():
        ev = net(x.to(cuda)).cpu()
    ref_ev = torch.from_numpy(g["eval_logits"])
    assert rel_err(ev, ref_ev) < TOL_LOGITS
    assert _argmax_agreement(ev, ref_ev, decided_only=True) >= MIN_AGREE
    assert _argmax_agreement(ev, ref_ev) >= 0.95  # all pixels, near-ties included


@pytest.mark.parametrize("name,n,h,w", [
    ("unet", 1, 45, 61),     # odd everywhere: floor pooling, F.pad on both axes at several levels
    ("unet", 3, 90, 120),    # 360x480 / 4: level-4 skip is 11 rows vs 10 upsampled (bottom pad), like 45 vs 44
    ("segnet", 1, 45, 70),   # unpool into odd output_size planes
    ("segnet", 3, 90, 120),
])
def test_train_step_matches_oracle(cvb, cuda, materialized, name, n, h, w):
    cutils, _ = cvb
    sd = _default_init_sd(cutils, name, 5)
    net = _build(cvb, name, sd, cuda)
    x, t = O.synth_batch(n, h, w, seed=6)
    loss, logits = _step(cvb, net, x, t, cuda)
    (o_loss, o_logits, o_grads, o_after, rec32), (m_loss, m_logits, m_grads, m_after, rec16) = \
        _oracle_steps_with_records(name, sd, x, t)
    _check_within_bf16_envelope(net, loss, logits, o_loss, o_logits, o_grads, m_logits, m_grads)
    _check_grad_norms(net, o_grads)
    _check_layer_trajectory(net, rec32, rec16, f"{name} {n}x{h}x{w}", logits)
    _check_logit_statistics(logits, o_logits, m_logits, f"{name} {n}x{h}x{w}")
    after = net.state_dict()
    stat_keys = [k for k in o_after if k.endswith(("running_mean", "running_var"))]
    for i, k in enumerate(stat_keys):
        # statistics of the first two blocks see (almost) exact inputs; deeper ones are statistics of activations that
        # are themselves several percent off (chaos), over as few as 8 pixels at the bottleneck of the small cases
        assert rel_err(after[k].cpu(), o_after[k]) < (1e-2 if i < 4 else 0.25), k


@pytest.mark.parametrize("name", ["unet", "segnet"])
def test_full_resolution_step(cvb, cuda, materialized, name):
    """BASELINE configs[0] geometry: batch 2 x 3 x 360 x 480, 12 classes, torch default init."""
    cutils, _ = cvb
    sd = _default_init_sd(cutils, name, 0)
    net = _build(cvb, name, sd, cuda)
    x, t = O.synth_batch(2, 360, 480, seed=0)
    loss, logits = _step(cvb, net, x, t, cuda)
    (o_loss, o_logits, o_grads, _, rec32), (m_loss, m_logits, m_grads, _, rec16) = \
        _oracle_steps_with_records(name, sd, x, t)
    _check_within_bf16_envelope(net, loss, logits, o_loss, o_logits, o_grads, m_logits, m_grads)
    _check_grad_norms(net, o_grads)
    _check_layer_trajectory(net, rec32, rec16, f"{name} 2x360x480", logits)
    _check_logit_statistics(logits, o_logits, m_logits, f"{name} 2x360x480")


@pytest.mark.parametrize("name,n,h,w", [("unet", 2, 90, 120), ("segnet", 2, 90, 120), ("unet", 1, 360, 480)])
def test_every_block_on_identical_inputs(cvb, cuda, name, n, h, w):
    """Teacher forcing: each conv+BN+ReLU block of the plan gets the fp32 oracle's own input activation and output
    gradient for that block. Its activation must meet 2e-2 against the fp32 oracle; its input / weight / BatchNorm
    gradients must meet 3e-2 against the bf16 storage model of the block and, against fp32, be no worse than that
    model (ReLU-mask flips from the bf16 conv output, see the module docstring)."""
    from camvid_b200 import ops
    cutils, _ = cvb
    sd = _default_init_sd(cutils, name, 7)
    net = _build(cvb, name, sd, cuda).train()
    x, t = O.synth_batch(n, h, w, seed=12)
    O.RECORD = {}
    try:
        _, _, o_grads, _ = O.train_step(name, sd, x, t)
        rec = O.RECORD
    finally:
        O.RECORD = None
    with torch.no_grad():
        net(x.to(cuda))  # builds the plan
    plan = _plan(net)
    names = {id(p): k for k, p in net.named_parameters()}
    flat = torch.zeros(plan.flat_size, device=cuda)
    worst = {}

    def upd(kind, key, e, limit):
        if e / limit > worst.get(kind, ("", 0.0, 1.0, 0.0))[3]:
            worst[kind] = (key, e, limit, e / limit)

    for bi, b in enumerate(plan.blocks):
        r = rec[b.name + ".conv" if b.name.startswith("upsample") else b.name]
        kw, kg, kb = names[id(b.conv.weight)], names[id(b.bn.weight)], names[id(b.bn.bias)]
        # bf16 storage model of this block on the same inputs (CPU, fp32 arithmetic)
        xm = r["x"].clone().requires_grad_(r["dx"] is not None)
        pm = [sd[k].clone().requires_grad_(True) for k in (kw, kg, kb)]
        a_m, _, _ = O._BlockBF16.apply(xm, pm[0], pm[1], pm[2], b.bn.eps)
        gm = torch.autograd.grad(a_m, pm + ([xm] if r["dx"] is not None else []), grad_outputs=O._r(r["dout"]))
        # the CUDA block
        if b.taps == 1:
            ops.im2col3x3(r["x"].to(cuda).contiguous(), b.x)
        else:
            b.x.zero_()
            b.x[..., :b.cin].copy_(r["x"].permute(0, 2, 3, 1).to(cuda))
        b.forward_train()
        act = b.a[..., :b.cout].float().permute(0, 3, 1, 2).cpu()
        upd("act_vs_fp32", b.name, rel_err(act, r["out"]), TOL_LOGITS)
        upd("act_vs_bf16_model", b.name, rel_err(act, a_m.detach()), TOL_LOGITS)
        da = torch.zeros(b.a.shape, dtype=torch.bfloat16, device=cuda)
        da[..., :b.cout].copy_(r["dout"].permute(0, 2, 3, 1).to(cuda))
        dx = torch.empty(b.x.shape, dtype=torch.bfloat16, device=cuda) if r["dx"] is not None else None
        b.backward(da, dx, flat)
        gw, _, gg, gb = plan.grads_for(flat)[4 * bi:4 * bi + 4]
        got = [gw.cpu(), gg.cpu(), gb.cpu()] + ([dx[..., :b.cin].float().permute(0, 3, 1, 2).cpu()] if dx is not None else [])
        ref32 = [o_grads[kw], o_grads[kg], o_grads[kb]] + ([r["dx"]] if dx is not None else [])
        for tag, g_cuda, g_model, g_fp32 in zip(("dw", "dgamma", "dbeta", "dx"), got, gm, ref32):
            upd(tag + "_vs_bf16_model", b.name, rel_err(g_cuda, g_model), TOL_GRAD)
            upd(tag + "_vs_fp32", b.name, rel_err(g_cuda, g_fp32), max(TOL_GRAD, 1.3 * rel_err(g_model, g_fp32)))
    torch.cuda.synchronize()
    print("worst per check (block, error, limit, ratio):", worst)
    for kind, (key, e, limit, ratio) in worst.items():
        assert ratio < 1.0, (kind, key, e, limit)


def test_ignore_index_void(cvb, cuda):
    """north_star asks for ignore_index (Void = 11) support: same step with CrossEntropyLoss(ignore_index=11)."""
    cutils, _ = cvb
    torch.manual_seed(1)
    sd = {k: v.clone() for k, v in cutils.get_model("unet", 3, 12).state_dict().items()}
    net = _build(cvb, "unet", sd, cuda)
    x, t = O.synth_batch(2, 48, 64, seed=8)
    loss, logits = _step(cvb, net, x, t, cuda, ignore_index=11)
    m_loss, m_logits, m_grads, _ = O.train_step("unet", sd, x, t, ignore_index=11, storage="bf16")
    o_loss, o_logits, o_grads, _ = O.train_step("unet", sd, x, t, ignore_index=11)
    _check_within_bf16_envelope(net, loss, logits, o_loss, o_logits, o_grads, m_logits, m_grads)
    _check_grad_norms(net, o_grads)


@pytest.mark.parametrize("fused_optimizer", [False, True])
def test_optimizer_steps_follow_oracle(cvb, cuda, fused_optimizer):
    """Three AdamW steps (train.py:100,124-134): packed bf16 weights are refreshed from the fp32 parameters after
    each optimizer.step(); the loss trajectory follows the fp32 oracle's -- with torch.optim.AdamW and with the fused
    drop-in camvid_b200.optim.AdamW (the oracle side always steps torch's AdamW on the CPU)."""
    cutils, cnn = cvb
    torch.manual_seed(2)
    sd = {k: v.clone() for k, v in cutils.get_model("segnet", 3, 12).state_dict().items()}
    net = _build(cvb, "segnet", sd, cuda)
    if fused_optimizer:
        from camvid_b200.optim import AdamW
    else:
        AdamW = torch.optim.AdamW
    opt = AdamW(net.parameters(), lr=5e-4, weight_decay=0)
    x, t = O.synth_batch(2, 64, 96, seed=9)
    # oracle side: functional model + the same optimizer on CPU tensors
    o_sd = {k: v.clone() for k, v in sd.items()}
    names = [k for k, _ in net.named_parameters()]
    o_params = [torch.nn.Parameter(o_sd[k].clone()) for k in names]
    o_opt = torch.optim.AdamW(o_params, lr=5e-4, weight_decay=0)
    losses, o_losses = [], []
    for _ in range(3):
        opt.zero_grad()
        loss = cnn.CrossEntropyLoss()(net.train()(x.to(cuda)), t.to(cuda))
        loss.backward()
        opt.step()
        losses.append(loss.item())
        for k, p in zip(names, o_params):
            o_sd[k] = p.detach().clone()
        o_loss, _, o_grads, o_after = O.train_step("segnet", o_sd, x, t)
        for k, p in zip(names, o_params):
            p.grad = o_grads[k]
        o_opt.step()
        o_sd = {k: v.clone() for k, v in o_after.items()}
        o_losses.append(o_loss.item())
    assert losses[2] < losses[0]  # it learns
    for a, b in zip(losses, o_losses):
        assert abs(a - b) / b < 3e-2, (losses, o_losses)


def test_stock_torch_cross_entropy_on_dropin_logits(cvb, cuda):
    """SURVEY 8b contract: the reference's own loss object, `nn.CrossEntropyLoss()` (train.py:105,130-131), applied to
    the drop-in module's logits, backward through torch's loss into the plan: same loss and gradients as with the fused
    camvid_b200.nn.CrossEntropyLoss."""
    cutils, cnn = cvb
    torch.manual_seed(4)
    sd = {k: v.clone() for k, v in cutils.get_model("unet", 3, 12).state_dict().items()}
    x, t = O.synth_batch(2, 48, 64, seed=14)
    results = []
    for loss_fn in (torch.nn.CrossEntropyLoss(), cnn.CrossEntropyLoss()):
        net = _build(cvb, "unet", sd, cuda).train()
        logits = net(x.to(cuda))
        loss = loss_fn(logits, t.to(cuda))
        loss.backward()
        results.append((loss.item(), logits.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    (l_t, lg_t, g_t), (l_c, lg_c, g_c) = results
    assert torch.equal(lg_t, lg_c)  # same kernels, same inputs
    assert abs(l_t - l_c) < 1e-5 * abs(l_t)
    o_loss, _, o_grads, _ = O.train_step("unet", sd, x, t)
    assert abs(l_t - o_loss.item()) / o_loss.item() < TOL_LOSS
    for k in g_t:
        if not _is_conv_bias(net, k):
            assert rel_err(g_c[k], g_t[k]) < 2e-3, k  # dlogits differ by fp32 rounding before the bf16 store


@pytest.mark.parametrize("fused_optimizer", [False, True])
def test_packed_weights_follow_the_optimizer(cvb, cuda, fused_optimizer):
    """the bf16 GEMM operands the conv kernels read are re-packed from the fp32 parameters after
    every optimizer step -- also with the fused AdamW, which writes the parameters through raw pointers. After
    step() + forward, every block's packed fprop / dgrad operand equals a fresh pack of its current fp32 weight, and the
    logits move although the BatchNorm affine parameters are frozen (only conv weights train)."""
    from camvid_b200 import ops
    cutils, cnn = cvb
    torch.manual_seed(6)
    net = cutils.get_model("segnet", 3, 12).to(cuda).train()
    conv_w = [p for k, p in net.named_parameters() if p.dim() == 4]
    if fused_optimizer:
        from camvid_b200.optim import AdamW
    else:
        AdamW = torch.optim.AdamW
    opt = AdamW(conv_w, lr=1e-2, weight_decay=0)  # BatchNorm gamma / beta and conv biases are NOT optimised
    x, t = O.synth_batch(2, 32, 48, seed=15)
    before = [w.detach().clone() for w in conv_w]
    logits0 = net(x.to(cuda))
    cnn.CrossEntropyLoss()(logits0, t.to(cuda)).backward()
    opt.step()
    assert all(not torch.equal(a, b) for a, b in zip(before, conv_w))
    with torch.no_grad():
        logits1 = net(x.to(cuda))
    plan = _plan(net)
    for b in plan.blocks:
        w = b.conv.weight.detach()
        assert torch.equal(b.wf, ops.pack_weights_fprop(w, b.taps, b.cout_pad, b.cin_pad)), b.name
        if b.wd is not None:
            assert torch.equal(b.wd, ops.pack_weights_dgrad(w, b.cout_pad, b.cin_pad)), b.name
    assert rel_err(logits1, logits0.detach()) > 1e-2  # lr 1e-2 on every conv weight: the output must move


def test_graphed_train_step_equals_eager_steps(cvb, cuda):
    """camvid_b200.graph.GraphedTrainStep: the whole step as one CUDA graph. Building it leaves parameters, running
    statistics and optimizer state untouched; N replays on N batches then equal N eager steps with the same optimizer --
    bit for bit (deterministic kernels, same launch sequence) -- including OneCycleLR changing lr / beta1 every step
    (train.py:102-104,134) and an eager eval forward afterwards seeing the updated weights."""
    from camvid_b200.graph import GraphedTrainStep
    from camvid_b200.optim import AdamW
    cutils, cnn = cvb
    torch.manual_seed(7)
    sd = {k: v.clone() for k, v in cutils.get_model("unet", 3, 12).state_dict().items()}
    batches = [O.synth_batch(2, 48, 64, seed=20 + i) for i in range(3)]
    batches = [(x.to(cuda), t.to(cuda)) for x, t in batches]
    results = []
    for graphed in (False, True):
        net = _build(cvb, "unet", sd, cuda).train()
        opt = AdamW(net.parameters(), lr=5e-4, weight_decay=1e-2, capturable=True)
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=5e-4, steps_per_epoch=3, epochs=2)
        loss_fn = cnn.CrossEntropyLoss()
        if graphed:
            step = GraphedTrainStep(net, loss_fn, opt, *batches[0])
            for k, v in net.state_dict().items():
                assert torch.equal(v.cpu(), sd[k]), k  # construction (warm-up + capture) changed nothing
            assert all(float(st["step"]) == 0.0 for st in opt.state.values())
        losses = []
        for x, t in batches:
            if graphed:
                losses.append(step(x, t).item())
            else:
                opt.zero_grad(set_to_none=True)
                loss = loss_fn(net(x), t)
                loss.backward()
                opt.step()
                losses.append(loss.item())
            sched.step()
        net.eval()
        with torch.no_grad():
            ev = net(batches[0][0]).clone()
        results.append((losses, {k: v.clone() for k, v in net.state_dict().items()}, ev))
    (l_e, sd_e, ev_e), (l_g, sd_g, ev_g) = results
    assert l_e == l_g, (l_e, l_g)
    for k in sd_e:
        assert torch.equal(sd_e[k], sd_g[k]), k
    assert torch.equal(ev_e, ev_g)
    assert l_e[2] != l_e[0]


def test_eval_metrics_pipeline(cvb, cuda):
    """eval.py:50-72 / train.py:180-197 on the device: logits -> argmax -> mean_iou + Metrics, equal to the oracle's
    metric functions applied to the same predictions (integer counts: bit-exact)."""
    cutils, cnn = cvb
    from camvid_b200.legacy.metrics import Metrics
    torch.manual_seed(3)
    net = cutils.get_model("unet", 3, 12).to(cuda).eval()
    x, t = O.synth_batch(2, 64, 96, seed=10)
    with torch.no_grad():
        logits = net(x.to(cuda))
        loss = cnn.CrossEntropyLoss()(logits, t.to(cuda))
    assert torch.isfinite(loss)
    preds = logits.argmax(dim=1)
    all_acc, acc, iou = cutils.mean_iou(preds, t.to(cuda), 12, 11)
    o_all, o_acc, o_iou = O.mean_iou(preds.cpu().numpy(), t.numpy(), 12, 11)
    assert all_acc == o_all
    np.testing.assert_array_equal(acc, o_acc)
    np.testing.assert_array_equal(iou, o_iou)
    m, om = Metrics(12, 11), O.Metrics(12, 11)
    m.add(preds.view(-1), t.to(cuda).view(-1))
    om.add(preds.view(-1).cpu().numpy(), t.view(-1).numpy())
    np.testing.assert_array_equal(m._confusion_matrix, om.cm)
    assert m.iou() == om.iou() and m.precision() == om.precision() and m.recall() == om.recall()
    m2 = Metrics(12, 11)
    m2.add_logits(logits, t.to(cuda))  # fused argmax + counting
    np.testing.assert_array_equal(m2._confusion_matrix, om.cm)


def test_several_forwards_before_a_backward(cvb, cuda):
    """A plan pool per input shape (engine.PLANS_PER_SHAPE = 2 sets of activation buffers): two recorded forwards can
    both be backpropagated, a no_grad validation forward between a training forward and its backward does not disturb
    it, and the gradients equal those of the same steps run one after the other. Beyond the pool the oldest pending
    forward is overwritten and its backward fails loudly; so does a second backward through the same forward."""
    from camvid_b200 import engine
    cutils, cnn = cvb
    torch.manual_seed(9)
    net = cutils.get_model("segnet", 3, 12).to(cuda).train()
    xa, ta = O.synth_batch(1, 32, 48, seed=11)
    xb, tb = O.synth_batch(1, 32, 48, seed=12)
    loss_fn = cnn.CrossEntropyLoss()

    def grads_of(x, t):
        net.zero_grad(set_to_none=True)
        loss_fn(net(x.to(cuda)), t.to(cuda)).backward()
        return {k: p.grad.clone() for k, p in net.named_parameters()}

    ga, gb = grads_of(xa, ta), grads_of(xb, tb)
    net.zero_grad(set_to_none=True)
    la = loss_fn(net(xa.to(cuda)), ta.to(cuda))
    with torch.no_grad():
        net(xb.to(cuda))  # validation-style forward in between: takes the pool's other plan, not the pending one
    lb = loss_fn(net(xb.to(cuda)), tb.to(cuda))  # second recorded forward before the first backward
    assert len(engine.plans_of(net)) == 2
    la.backward()
    lb.backward()
    for k, p in net.named_parameters():
        assert torch.equal(p.grad, ga[k] + gb[k]), k  # deterministic kernels: bit-identical to the sequential steps
    # a released graph frees its plan: many forwards in a row never grow the pool
    for _ in range(4):
        loss_fn(net(xa.to(cuda)), ta.to(cuda))
    assert len(engine.plans_of(net)) == 2
    # three live graphs on a pool of two: the oldest is overwritten and says so
    l1 = loss_fn(net(xa.to(cuda)), ta.to(cuda))
    l2 = loss_fn(net(xa.to(cuda)), ta.to(cuda))
    l3 = loss_fn(net(xa.to(cuda)), ta.to(cuda))
    with pytest.raises(RuntimeError, match="overwritten"):
        l1.backward()
    l2.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        l2.backward()  # (the fused loss says so first; below, the plan itself through a plain torch reduction)
    l3.backward()
    s = net(xa.to(cuda)).sum()
    s.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        s.backward()


def test_plan_cache_evicts_old_geometries(cvb, cuda):
    """module._plans keeps engine.MAX_SHAPES input geometries (least recently used dropped): a ragged last batch or a
    sweep over image sizes does not pin one set of buffers per shape forever."""
    from camvid_b200 import engine
    cutils, _ = cvb
    net = cutils.get_model("unet", 3, 12).to(cuda).eval()
    with torch.no_grad():
        for i in range(engine.MAX_SHAPES + 2):
            net(torch.zeros(1, 3, 32, 32 + 16 * i, device=cuda))
        shapes = list(net.__dict__["_plans"])
        assert len(shapes) == engine.MAX_SHAPES and shapes[-1][:3] == (1, 32, 32 + 16 * (engine.MAX_SHAPES + 1))
        net(torch.zeros(1, 3, 32, 32 + 16 * 2, device=cuda))  # re-use moves a geometry to the back of the queue
        assert list(net.__dict__["_plans"])[-1][:3] == (1, 32, 64)


@pytest.mark.parametrize("name", ["unet", "segnet"])
def test_sublayers_have_their_own_forward(cvb, cuda, name):
    """VERDICT r1 / north_star: "their layers dispatch through a thin C-ABI torch custom-op layer". BasicConv2d,
    UpSample2d (models/unet.py:16-17,28-32) and BasicConv (models/segnet.py:14-17) used directly -- `net.down1(x)`, a
    hand-written forward over the sub-modules -- run the same kernels, differentiable w.r.t. input and parameters,
    against the same layer of the fp32 reference arithmetic (torch ops on the CPU)."""
    import torch.nn.functional as F
    cutils, _ = cvb
    torch.manual_seed(13)
    net = cutils.get_model(name, 3, 12).to(cuda).train()
    if name == "unet":
        stage, layer_names = net.down2, ["down2.0", "down2.1"]
        convs = [(b.conv[0], b.conv[1]) for b in stage]
        cin = 64
    else:
        stage, layer_names = net.encoder2, ["encoder2.0", "encoder2.1"]
        convs = [(b.conv, b.bn) for b in stage]
        cin = 64
    x = torch.relu(torch.randn(2, cin, 20, 28)).to(torch.bfloat16).float()
    xg = x.to(cuda).requires_grad_(True)
    out = stage(xg)  # nn.Sequential over two sub-layers, each a block_forward op
    assert tuple(out.shape) == (2, 128, 20, 28) and out.requires_grad
    out.backward(torch.randn_like(out))
    assert torch.isfinite(xg.grad).all() and all(int(bn.num_batches_tracked) == 1 for _, bn in convs)
    # each layer on its own against the fp32 block and its bf16 storage model (oracle.block_step), like the in-plan
    # teacher-forced block tests
    inp = x
    for layer, (conv, bn) in zip(stage, convs):
        for p_ in layer.parameters():
            p_.grad = None
        dout = torch.randn(2, conv.out_channels, 20, 28).to(torch.bfloat16).float()
        wt, bias = conv.weight.detach().cpu(), conv.bias.detach().cpu()
        gamma, beta = bn.weight.detach().cpu(), bn.bias.detach().cpu()
        r32 = O.block_step(inp, wt, bias, gamma, beta, dout)
        r16 = O.block_step(inp, wt, bias, gamma, beta, dout, storage="bf16")
        xi = inp.to(cuda).requires_grad_(True)
        a = layer(xi)
        a.backward(dout.to(cuda))
        got = (a.detach().cpu(), xi.grad.cpu(), conv.weight.grad.cpu(), bn.weight.grad.cpu(), bn.bias.grad.cpu())
        assert conv.bias.grad is not None and conv.bias.grad.abs().max().item() < 1e-4
        assert rel_err(got[0], r32[0]) < TOL_LOGITS
        for g_cuda, g16, g32 in zip(got[1:], r16[1:], r32[1:]):
            assert rel_err(g_cuda, g16) < TOL_GRAD
            assert rel_err(g_cuda, g32) < max(TOL_GRAD, 1.3 * rel_err(g16, g32))
        inp = r16[0]  # a bf16-representable activation as the next layer's input
    if name == "unet":  # UpSample2d: bilinear x2 (align_corners) + block, eval mode too
        up = net.upsample4
        xu = torch.relu(torch.randn(1, 128, 9, 11)).to(cuda).requires_grad_(True)
        y = up(xu)
        assert tuple(y.shape) == (1, 64, 18, 22)
        conv, bn = up.conv.conv[0], up.conv.conv[1]
        xc = xu.detach().cpu().requires_grad_(True)
        r = F.interpolate(xc, scale_factor=2, mode="bilinear", align_corners=True)
        r = F.relu(F.batch_norm(F.conv2d(r, conv.weight.detach().cpu(), conv.bias.detach().cpu(), padding=1), None, None,
                                bn.weight.detach().cpu(), bn.bias.detach().cpu(), True, 0.1, bn.eps))
        g = torch.randn_like(r)
        y.backward(g.to(cuda))
        r.backward(g)
        assert rel_err(y.detach().cpu(), r.detach()) < TOL_LOGITS
        assert rel_err(xu.grad.cpu(), xc.grad) < 2 * TOL_GRAD  # bf16 upsampled operand + ReLU-mask flips of one block
        up.eval()
        with torch.no_grad():
            ye = up(xu.detach())
            re_ = F.relu(F.batch_norm(F.conv2d(F.interpolate(xc.detach(), scale_factor=2, mode="bilinear", align_corners=True),
                                               conv.weight.detach().cpu(), conv.bias.detach().cpu(), padding=1),
                                      bn.running_mean.cpu(), bn.running_var.cpu(), bn.weight.detach().cpu(),
                                      bn.bias.detach().cpu(), False, 0.1, bn.eps))
        assert rel_err(ye.cpu(), re_) < TOL_LOGITS


@pytest.mark.parametrize("name,in_ch", [("unet", 8), ("segnet", 20), ("unet", 1)])
def test_input_channel_counts_other_than_camvids(cvb, cuda, name, in_ch):
    """utils.get_model(name, input_channels, class_num) (utils.py:147-160) takes any channel count: up to 7 channels go
    through the first-layer im2col, wider inputs through the ordinary 9-tap path. Eval mode against the fp32 oracle at
    the north_star tolerance, train mode: loss and first-layer gradient norm."""
    cutils, cnn = cvb
    torch.manual_seed(41)
    net = cutils.get_model(name, in_ch, 5)
    sd = O.synth_state_dict(net.state_dict(), seed=3)
    net.load_state_dict(sd)
    net = net.to(cuda)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, in_ch, 48, 64, generator=g)
    t = torch.randint(0, 5, (2, 48, 64), generator=g)
    net.eval()
    with torch.no_grad():
        ev = net(x.to(cuda)).cpu()
    ref = O.forward(name, sd, x, train=False)
    assert rel_err(ev, ref) < TOL_LOGITS
    net.train()
    loss = cnn.CrossEntropyLoss()(net(x.to(cuda)), t.to(cuda))
    loss.backward()
    o_loss, _, o_grads, _ = O.train_step(name, sd, x, t)
    assert abs(loss.item() - float(o_loss)) < TOL_LOGITS * abs(float(o_loss))
    first = next(k for k, v in o_grads.items() if v.dim() == 4)  # the first convolution's weight
    gw = dict(net.named_parameters())[first].grad.cpu()
    assert tuple(gw.shape)[1] == in_ch
    assert abs(gw.norm().item() / o_grads[first].norm().item() - 1.0) < 0.15


def test_jit_trace_records_one_dispatcher_op(cvb, cuda):
    """utils.visualize_network (utils.py:10-13; train.py:97-98) traces the TRAIN-mode module with torch.jit.trace on a
    [1,3,480,360]-shaped tensor (IMAGE_SIZE transposed). The drop-in must trace: the network is one
    `camvid_b200::net_forward` node, and the traced module reproduces the eager output."""
    cutils, _ = cvb
    net = cutils.get_model("unet", 3, 12).to(cuda).train()
    x = torch.randn(1, 3, 48, 36, device=cuda)
    traced = torch.jit.trace(net, x, strict=False)
    assert "camvid_b200::net_forward" in str(traced.inlined_graph)
    assert torch.equal(traced(x), net(x))


def test_custom_op_schema_and_eval_backward_guard(cvb, cuda):
    cutils, cnn = cvb
    assert "Tensor[] params" in str(torch.ops.camvid_b200.net_forward.default._schema)
    net = cutils.get_model("segnet", 3, 12).to(cuda).eval()
    x, t = O.synth_batch(1, 32, 32, seed=12)
    loss = cnn.CrossEntropyLoss()(net(x.to(cuda)), t.to(cuda))
    with pytest.raises(RuntimeError, match="eval-mode"):
        loss.backward()


def test_fused_bn_backward_statistics_path(cvb, cuda):
    """engine.FUSE_BWD_STATS (off by default: slower end to end) routes the BatchNorm backward reduction of 64-channel
    blocks through the data-gradient epilogue (cvb_conv_epilogue.bwd_*). Same sums from the same stored tensors: the
    parameter gradients must agree with the default path to accumulation-order noise."""
    from camvid_b200 import engine
    cutils, cnn = cvb
    grads = []
    for fuse in (False, True):
        engine.FUSE_BWD_STATS = fuse
        try:
            torch.manual_seed(5)
            net = cutils.get_model("segnet", 3, 12).to(cuda).train()
            x, t = O.synth_batch(2, 64, 48, seed=13)
            cnn.CrossEntropyLoss()(net(x.to(cuda)), t.to(cuda)).backward()
            grads.append({k: p.grad.clone() for k, p in net.named_parameters()})
        finally:
            engine.FUSE_BWD_STATS = False
    worst = max(rel_err(grads[1][k], grads[0][k]) for k in grads[0] if not _is_conv_bias(net, k))
    assert worst < 2e-3, worst
