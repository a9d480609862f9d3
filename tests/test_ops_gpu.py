"""Op-level parity of the CUDA kernels (through the C ABI) against the ATen ops the reference calls.

Oracle = torch fp32 ops on the same (bf16-representable) inputs, TF32 disabled. Integer / index results must be
bit-exact; floating-point results are compared norm-wise with the tolerance stated in each test.
"""
import pytest
import torch
import torch.nn.functional as F

from util import bf16_round, rel_err, to_nchw_f32, to_nhwc_bf16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    import camvid_b200  # noqa: F401
    from camvid_b200 import ops as _ops
    return _ops


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return bf16_round(torch.randn(*shape, generator=g) * scale).to(dev)


# ------------------------------------------------------------------------------------------- layout
def test_layout_roundtrip(ops, cuda):
    x = _rand((2, 3, 13, 17), cuda, 0)
    dst = torch.full((2, 13, 17, 64), 7.0, dtype=torch.bfloat16, device=cuda)
    ops.nchw_to_nhwc(x, dst)
    assert torch.equal(dst[..., :3].float(), x.permute(0, 2, 3, 1))
    assert torch.count_nonzero(dst[..., 3:]) == 0
    back = torch.empty(2, 3, 13, 17, device=cuda)
    ops.nhwc_to_nchw(dst, back)
    assert torch.equal(back, x)


def test_im2col(ops, cuda):
    x = _rand((2, 3, 9, 11), cuda, 1)
    dst = torch.empty(2, 9, 11, 64, dtype=torch.bfloat16, device=cuda)
    ops.im2col3x3(x, dst)
    ref = F.unfold(x, 3, padding=1).view(2, 27, 9, 11).permute(0, 2, 3, 1)  # k = ci*9 + r*3 + s
    assert torch.equal(dst[..., :27].float(), ref)
    assert torch.count_nonzero(dst[..., 27:]) == 0


# ------------------------------------------------------------------------------------------- convolution
CONV_CASES = [
    # n, h, w, cin, cout
    (2, 16, 24, 64, 64),
    (1, 45, 60, 128, 256),
    (3, 22, 30, 64, 512),
    (2, 11, 15, 192, 64),
    (1, 8, 136, 64, 128),
    (16, 5, 7, 128, 128),
    # halo kernel (cout <= 128, h >= 16): ragged right / bottom edges, skipped lower half, several cin chunks
    (1, 45, 61, 64, 64),
    (2, 33, 20, 128, 128),
    (1, 70, 9, 256, 64),
    (3, 17, 8, 64, 128),
    (1, 96, 40, 192, 128),
    # CTA-pair halo kernel: more tile pairs than SM pairs (several accumulator phases), odd tile count
    (8, 128, 100, 64, 64),
    (5, 96, 88, 128, 128),
]


@pytest.mark.parametrize("n,h,w,cin,cout", CONV_CASES)
def test_conv_fprop(ops, cuda, n, h, w, cin, cout):
    x = _rand((n, cin, h, w), cuda, 2)
    wt = _rand((cout, cin, 3, 3), cuda, 3, scale=(cin * 9) ** -0.5)
    ref = F.conv2d(x, wt, padding=1)
    xs = to_nhwc_bf16(x)
    wp = ops.pack_weights_fprop(wt, 9, cout, cin)
    y = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=cuda)
    parts = torch.full((ops.stat_rows(), 2, cout), float("nan"), device=cuda)
    ops.conv3x3(xs, wp, y, stat_partials=parts)
    torch.cuda.synchronize()
    got = to_nchw_f32(y)
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) < 4e-3  # bf16 output rounding only (operands are exact in bf16, fp32 accumulate)
    # fused BatchNorm statistics: per-channel sum and sum of squares of the fp32 accumulators
    s = parts.sum(0)
    assert rel_err(s[0], ref.sum((0, 2, 3))) < 2e-3
    assert rel_err(s[1], (ref * ref).sum((0, 2, 3))) < 2e-3


def test_conv_fprop_eval_epilogue_and_strided_output(ops, cuda):
    n, h, w, cin, cout = 2, 12, 20, 64, 64
    x = _rand((n, cin, h, w), cuda, 4)
    wt = _rand((cout, cin, 3, 3), cuda, 5, scale=(cin * 9) ** -0.5)
    scale = torch.rand(cout, device=cuda) + 0.5
    shift = torch.randn(cout, device=cuda) * 0.3
    ref = F.relu(F.conv2d(x, wt, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    big = torch.full((n, h + 1, w, 192), 3.0, dtype=torch.bfloat16, device=cuda)
    y = big[:, :h, :, 64:128]  # channel slice of a concat buffer with one extra (pad) row
    ops.conv3x3(to_nhwc_bf16(x), ops.pack_weights_fprop(wt, 9, cout, cin), y, scale=scale, shift=shift, relu=True)
    assert rel_err(to_nchw_f32(y), ref) < 4e-3
    assert (big[:, h] == 3).all() and (big[..., :64] == 3).all() and (big[..., 128:] == 3).all()


@pytest.mark.parametrize("cols_c", [64, 32])  # 32: narrow im2col buffer, the rest of the K chunk is TMA zero fill
def test_conv_first_layer_im2col(ops, cuda, cols_c):
    n, h, w, cin, cout = 2, 20, 28, 3, 64
    x = _rand((n, cin, h, w), cuda, 6)
    wt = _rand((cout, cin, 3, 3), cuda, 7, scale=27 ** -0.5)
    ref = F.conv2d(x, wt, padding=1)
    cols = torch.empty(n, h, w, cols_c, dtype=torch.bfloat16, device=cuda)
    ops.im2col3x3(x, cols)
    wp = ops.pack_weights_fprop(wt, 1, 64, 64)
    y = torch.empty(n, h, w, 64, dtype=torch.bfloat16, device=cuda)
    ops.conv3x3(cols, wp, y, taps=1)
    assert rel_err(to_nchw_f32(y), ref) < 4e-3


@pytest.mark.parametrize("n,h,w,cin,cout", CONV_CASES[:4])
def test_conv_dgrad(ops, cuda, n, h, w, cin, cout):
    x = _rand((n, cin, h, w), cuda, 8).requires_grad_(True)
    wt = _rand((cout, cin, 3, 3), cuda, 9, scale=(cin * 9) ** -0.5)
    dy = _rand((n, cout, h, w), cuda, 10)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), x, dy)
    wp = ops.pack_weights_dgrad(wt, cout, cin)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=cuda)
    ops.conv3x3(to_nhwc_bf16(dy), wp, dx)
    assert rel_err(to_nchw_f32(dx), ref) < 4e-3



def test_conv_cta_pair_kernel_matches_single_cta(ops, cuda, monkeypatch):
    """conv_fprop_halo2_kernel (cta_group::2, off by default because it is slower on B200) stays bit-identical to the
    single-CTA halo kernel: same MMAs, same accumulation order."""
    n, h, w, cin, cout = 3, 50, 44, 128, 128
    x = to_nhwc_bf16(_rand((n, cin, h, w), cuda, 41))
    wp = ops.pack_weights_fprop(_rand((cout, cin, 3, 3), cuda, 42, scale=(cin * 9) ** -0.5), 9, cout, cin)
    outs, stats = [], []
    monkeypatch.setenv("CVB_TR128", "0")  # cout = 128 normally takes the transposed kernel
    for mode in ("0", "1"):
        monkeypatch.setenv("CVB_HALO_PAIR", mode)
        y = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=cuda)
        parts = torch.full((ops.stat_rows(), 2, cout), float("nan"), device=cuda)
        ops.conv3x3(x, wp, y, stat_partials=parts)
        torch.cuda.synchronize()
        outs.append(y)
        stats.append(parts.sum(0))
    assert torch.equal(outs[0], outs[1])
    assert rel_err(stats[1], stats[0]) < 1e-5


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 40, 24, 64, 64), (1, 70, 20, 64, 128), (3, 128, 100, 64, 64)])
def test_conv_dgrad_emits_consumer_bn_backward_statistics(ops, cuda, n, h, w, cin, cout):
    """cvb_conv_epilogue.bwd_*: the data-gradient kernel reduces g = da*[y*scale+shift > 0] and g*y for the block that
    consumes da; must equal cvb_bn_relu_bwd_reduce run on the stored da (aten::native_batch_norm_backward's sums)."""
    wt = _rand((cout, cin, 3, 3), cuda, 31, scale=(cin * 9) ** -0.5)
    dy = to_nhwc_bf16(_rand((n, cout, h, w), cuda, 32))
    y_prev = to_nhwc_bf16(_rand((n, cin, h, w), cuda, 33))
    scale = torch.rand(cin, device=cuda) + 0.5
    shift = torch.randn(cin, device=cuda) * 0.3
    wd = ops.pack_weights_dgrad(wt, cout, cin)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=cuda)
    assert ops.conv3x3_fuses_bwd_stats(dy, dx)
    parts = torch.full((ops.stat_rows(), 2, cin), float("nan"), device=cuda)
    ops.conv3x3(dy, wd, dx, bwd=(y_prev, scale, shift, parts))
    dx_plain = torch.empty_like(dx)
    ops.conv3x3(dy, wd, dx_plain)
    assert torch.equal(dx, dx_plain)
    rows = 4 * ops.sm_count()
    ref = torch.empty(rows, 2, cin, device=cuda)
    ops.bn_relu_bwd_reduce(dx, y_prev, scale, shift, ref, rows)
    got, want = parts.double().sum(0), ref.double().sum(0)
    assert torch.isfinite(got).all()
    assert rel_err(got[0], want[0]) < 1e-4 and rel_err(got[1], want[1]) < 1e-4



def test_conv_dgrad_narrow_dy_skips_zero_k_steps(ops, cuda):
    """Output layer (12 classes padded to 64 channels of dy): reading dy through its first 16 channels only (the rest is
    TMA zero fill, the all-zero K steps are skipped) must give the same data gradient as the padded view."""
    n, h, w, cin, cout = 2, 36, 40, 64, 12
    wt = _rand((cout, cin, 3, 3), cuda, 51, scale=(cin * 9) ** -0.5)
    dyp = torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device=cuda)
    dyp[..., :cout] = to_nhwc_bf16(_rand((n, cout, h, w), cuda, 52))
    wd = ops.pack_weights_dgrad(wt, 64, cin)
    full, narrow = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=cuda), torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=cuda)
    ops.conv3x3(dyp, wd, full)
    ops.conv3x3(dyp[..., :16], wd, narrow)
    assert torch.equal(full, narrow)


WGRAD_CASES = [
    (2, 16, 24, 64, 64),
    (2, 23, 31, 128, 64),
    (2, 16, 16, 64, 128),
    (1, 45, 60, 128, 256),
    (3, 22, 30, 256, 128),
    (2, 11, 15, 512, 512),
    # ragged tiles (8 x 16 pixel tiles), odd row counts (half-filled last K slice), stream-K ranges crossing items
    (1, 45, 61, 64, 64),
    (1, 17, 9, 192, 64),
    (5, 8, 8, 64, 256),
    (1, 33, 70, 1024, 64),
]


@pytest.mark.parametrize("n,h,w,cin,cout", WGRAD_CASES)
def test_conv_wgrad(ops, cuda, n, h, w, cin, cout):
    x = _rand((n, cin, h, w), cuda, 11)
    wt = _rand((cout, cin, 3, 3), cuda, 12).requires_grad_(True)
    dy = _rand((n, cout, h, w), cuda, 13)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), wt, dy)
    dw = torch.full((cout, cin, 3, 3), float("nan"), device=cuda)
    ops.conv3x3_wgrad(to_nhwc_bf16(x), to_nhwc_bf16(dy), dw)
    assert torch.isfinite(dw).all()
    assert rel_err(dw, ref) < 1e-4  # fp32 accumulate of exact bf16 products, fp32 output


@pytest.mark.parametrize("cols_c", [64, 32])
def test_conv_wgrad_first_layer_and_padded_cout(ops, cuda, cols_c):
    n, h, w = 2, 20, 28
    x = _rand((n, 3, h, w), cuda, 14)
    wt = _rand((64, 3, 3, 3), cuda, 15).requires_grad_(True)
    dy = _rand((n, 64, h, w), cuda, 16)
    (ref,) = torch.autograd.grad(F.conv2d(x, wt, padding=1), wt, dy)
    cols = torch.empty(n, h, w, cols_c, dtype=torch.bfloat16, device=cuda)
    ops.im2col3x3(x, cols)
    dw = torch.empty(64, 3, 3, 3, device=cuda)
    ops.conv3x3_wgrad(cols, to_nhwc_bf16(dy), dw, taps=1)
    assert rel_err(dw, ref) < 1e-4
    # last layer: cout = 12 padded to 64 channels of dy (pad channels zero)
    x2 = _rand((n, 64, h, w), cuda, 17)
    w2 = _rand((12, 64, 3, 3), cuda, 18).requires_grad_(True)
    dy2 = _rand((n, 12, h, w), cuda, 19)
    (ref2,) = torch.autograd.grad(F.conv2d(x2, w2, padding=1), w2, dy2)
    dyp = torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device=cuda)
    dyp[..., :12] = to_nhwc_bf16(dy2)
    dw2 = torch.empty(12, 64, 3, 3, device=cuda)
    ops.conv3x3_wgrad(to_nhwc_bf16(x2), dyp, dw2)
    assert rel_err(dw2, ref2) < 1e-4


# ------------------------------------------------------------------------------------------- batch norm
@pytest.mark.parametrize("n,h,w,c,creal", [(2, 9, 13, 64, 64), (3, 8, 8, 256, 256), (2, 10, 6, 64, 12), (1, 7, 5, 1024, 1024)])
def test_bn_relu_train(ops, cuda, n, h, w, c, creal):
    torch.manual_seed(20)
    y = bf16_round(torch.randn(n, c, h, w) * 1.7 + 0.4).to(cuda)
    y[:, creal:] = 0
    gamma = (torch.rand(creal) + 0.5).to(cuda)
    beta = (torch.randn(creal) * 0.2).to(cuda)
    bias = (torch.randn(creal) * 0.1).to(cuda)
    rm = torch.randn(creal).to(cuda)
    rv = (torch.rand(creal) + 0.5).to(cuda)
    da = bf16_round(torch.randn(n, c, h, w)).to(cuda)
    # oracle: F.batch_norm (train) + relu on (y + bias); the kernels drop the bias and add it back to running_mean
    yr = (y[:, :creal] + bias.view(1, -1, 1, 1)).clone().requires_grad_(True)
    g_ = gamma.clone().requires_grad_(True)
    b_ = beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    a_ref = F.relu(F.batch_norm(yr, rm_ref, rv_ref, g_, b_, True, 0.1, 1e-5))
    dy_ref, dg_ref, db_ref = torch.autograd.grad(a_ref, (yr, g_, b_), da[:, :creal])

    ys, das = to_nhwc_bf16(y), to_nhwc_bf16(da)
    rows = 37
    parts = torch.empty(rows, 2, c, device=cuda)
    ops.bn_stats(ys, parts, rows)
    mean, invstd, scale, shift = (torch.empty(c, device=cuda) for _ in range(4))
    count = n * h * w
    ops.bn_finalize(parts, rows, creal, c, count, gamma, beta, bias, rm, rv, 0.1, 1e-5, mean, invstd, scale, shift)
    a = torch.empty_like(ys)
    ops.bn_relu_apply(ys, scale, shift, a)
    assert rel_err(to_nchw_f32(a)[:, :creal], a_ref) < 4e-3
    assert torch.count_nonzero(a[..., creal:]) == 0
    assert torch.allclose(rm, rm_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-6)
    # backward
    ops.bn_relu_bwd_reduce(das, ys, scale, shift, parts, rows)
    dgamma, dbeta = torch.empty(creal, device=cuda), torch.empty(creal, device=cuda)
    coef = torch.empty(3, c, device=cuda)
    ops.bn_bwd_finalize(parts, rows, creal, c, count, gamma, mean, invstd, dgamma, dbeta, coef)
    dy = torch.empty_like(ys)
    ops.bn_relu_bwd_apply(das, ys, scale, shift, coef, dy)
    assert rel_err(dgamma, dg_ref) < 1e-3
    assert rel_err(dbeta, db_ref) < 1e-3
    assert rel_err(to_nchw_f32(dy)[:, :creal], dy_ref) < 5e-3
    assert torch.count_nonzero(dy[..., creal:]) == 0


@pytest.mark.parametrize("n,h,w,cmem,stride_c,creal", [(2, 36, 40, 16, 16, 12), (1, 45, 61, 16, 64, 12),
                                                        (3, 17, 9, 8, 8, 5), (2, 360, 480, 16, 16, 12)])
def test_last_block_fused_with_the_module_boundary_is_bit_identical(ops, cuda, n, h, w, cmem, stride_c, creal):
    """cvb_bn_relu_apply_nchw_f32 == cvb_bn_relu_apply + cvb_nhwc_bf16_to_nchw_f32 and
    cvb_nchw_f32_to_nhwc_bf16_bn_reduce == cvb_nchw_f32_to_nhwc_bf16 + cvb_bn_relu_bwd_reduce (same da bits, same sums
    up to fp32 summation order), on dense and channel-strided views (the odd-height plans keep 64 channels in memory)."""
    buf = torch.zeros(n, h, w, stride_c, dtype=torch.bfloat16, device=cuda)
    buf[..., :creal] = to_nhwc_bf16(_rand((n, creal, h, w), cuda, 61))
    y = buf[..., :cmem]
    scale = torch.zeros(cmem, device=cuda)
    shift = torch.zeros(cmem, device=cuda)
    scale[:creal] = torch.rand(creal, device=cuda) + 0.5
    shift[:creal] = torch.randn(creal, device=cuda) * 0.3
    # forward
    a = torch.empty(n, h, w, cmem, dtype=torch.bfloat16, device=cuda)
    want = torch.empty(n, creal, h, w, device=cuda)
    ops.bn_relu_apply(y, scale, shift, a)
    ops.nhwc_to_nchw(a, want)
    got = torch.full((n, creal, h, w), float("nan"), device=cuda)
    ops.bn_relu_apply_nchw(y, scale, shift, got)
    assert torch.equal(got, want)
    # backward entry
    dl = torch.randn(n, creal, h, w, device=cuda) * 1e-3  # fp32, not bf16-representable: the kernel rounds
    rows = 4 * ops.sm_count()
    da_want = torch.full((n, h, w, cmem), 7.0, dtype=torch.bfloat16, device=cuda)
    ref = torch.empty(rows, 2, cmem, device=cuda)
    ops.nchw_to_nhwc(dl, da_want)
    ops.bn_relu_bwd_reduce(da_want, y, scale, shift, ref, rows)
    da_buf = torch.full((n, h, w, stride_c), 7.0, dtype=torch.bfloat16, device=cuda)
    da_got = da_buf[..., :cmem]
    parts = torch.full((rows, 2, cmem), float("nan"), device=cuda)
    ops.nchw_to_nhwc_bn_reduce(dl, da_got, y, scale, shift, parts, rows)
    assert torch.equal(da_got, da_want)
    if stride_c > cmem:
        assert (da_buf[..., cmem:] == 7.0).all()  # nothing written outside the view
    g, r = parts.double().sum(0), ref.double().sum(0)
    assert torch.isfinite(g).all()
    assert rel_err(g[0], r[0]) < 1e-5 and rel_err(g[1], r[1]) < 1e-5


# ------------------------------------------------------------------------------------------- pooling
@pytest.mark.parametrize("n,h,w,c", [(2, 8, 12, 64), (1, 45, 61, 64), (3, 11, 15, 128)])
def test_maxpool_indices_bit_exact(ops, cuda, n, h, w, c):
    torch.manual_seed(21)
    x = F.relu(bf16_round(torch.randn(n, c, h, w))).to(cuda)  # ~half zeros -> plenty of ties
    x[0, 0, 0, 1] = float("nan")
    x[0, 1, 1, 0] = float("nan")
    x[0, 1, 1, 1] = float("nan")
    ref, idx_ref = F.max_pool2d(x, 2, return_indices=True)
    xs = to_nhwc_bf16(x)
    out = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=cuda)
    code = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=cuda)
    ops.maxpool2x2(xs, out, code)
    got = to_nchw_f32(out)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    assert torch.equal(ops.pool_code_to_index(code, w), idx_ref)
    # backward = scatter; must equal torch's max_pool2d backward exactly (values are copied, not computed)
    dout = bf16_round(torch.randn(n, c, h // 2, w // 2)).to(cuda)
    dx_ref = F.max_unpool2d(dout, idx_ref, 2, output_size=x.shape)  # aten's max_pool2d backward is this scatter
    dx = torch.full((n, h, w, c), 5.0, dtype=torch.bfloat16, device=cuda)
    ops.maxpool2x2_bwd(to_nhwc_bf16(dout), dx, code=code)
    assert torch.equal(to_nchw_f32(dx), dx_ref)
    # recompute-from-input variant and accumulate variant
    dx2 = torch.ones(n, h, w, c, dtype=torch.bfloat16, device=cuda)
    ops.maxpool2x2_bwd(to_nhwc_bf16(dout), dx2, x=xs, accumulate=True)
    assert torch.equal(to_nchw_f32(dx2), bf16_round(dx_ref + 1.0))


@pytest.mark.parametrize("n,h,w,c", [(2, 8, 12, 64), (1, 45, 61, 64)])
def test_maxunpool_bit_exact(ops, cuda, n, h, w, c):
    torch.manual_seed(22)
    x = F.relu(bf16_round(torch.randn(n, c, h, w))).to(cuda)
    pooled, idx = F.max_pool2d(x, 2, return_indices=True)
    v = bf16_round(torch.randn_like(pooled))
    ref = F.max_unpool2d(v, idx, 2, output_size=x.shape)
    code = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=cuda)
    tmp = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=cuda)
    ops.maxpool2x2(to_nhwc_bf16(x), tmp, code)
    out = torch.full((n, h, w, c), 9.0, dtype=torch.bfloat16, device=cuda)
    ops.maxunpool2x2(to_nhwc_bf16(v), code, out)
    assert torch.equal(to_nchw_f32(out), ref)
    dout = bf16_round(torch.randn(n, c, h, w)).to(cuda)
    vr = v.clone().requires_grad_(True)
    (dv_ref,) = torch.autograd.grad(F.max_unpool2d(vr, idx, 2, output_size=x.shape), vr, dout)
    dv = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=cuda)
    ops.maxunpool2x2_bwd(to_nhwc_bf16(dout), code, dv)
    assert torch.equal(to_nchw_f32(dv), dv_ref)


def test_bn_relu_maxpool_fused(ops, cuda):
    n, h, w, c = 2, 9, 14, 64
    torch.manual_seed(23)
    y = to_nhwc_bf16(bf16_round(torch.randn(n, c, h, w)).to(cuda))
    scale = (torch.rand(c) + 0.5).to(cuda)
    shift = (torch.randn(c) * 0.3).to(cuda)
    a_ref = torch.empty_like(y)
    ops.bn_relu_apply(y, scale, shift, a_ref)
    p_ref = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=cuda)
    c_ref = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=cuda)
    ops.maxpool2x2(a_ref, p_ref, c_ref)
    a, p, cd = torch.empty_like(a_ref), torch.empty_like(p_ref), torch.empty_like(c_ref)
    ops.bn_relu_maxpool2x2(y, scale, shift, a, p, cd)
    assert torch.equal(a, a_ref) and torch.equal(p, p_ref) and torch.equal(cd, c_ref)


# ------------------------------------------------------------------------------------------- upsample
@pytest.mark.parametrize("n,h,w,c", [(2, 5, 7, 64), (1, 22, 30, 128), (1, 1, 3, 64)])
def test_bilinear2x(ops, cuda, n, h, w, c):
    torch.manual_seed(24)
    x = bf16_round(torch.randn(n, c, h, w)).to(cuda).requires_grad_(True)
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    out = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=cuda)
    ops.bilinear2x(to_nhwc_bf16(x.detach()), out)
    assert rel_err(to_nchw_f32(out), ref) < 4e-3
    dout = bf16_round(torch.randn(n, c, 2 * h, 2 * w)).to(cuda)
    (dx_ref,) = torch.autograd.grad(ref, x, dout)
    dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=cuda)
    ops.bilinear2x_bwd(to_nhwc_bf16(dout), dx)
    assert rel_err(to_nchw_f32(dx), dx_ref) < 4e-3


@pytest.mark.parametrize("n,h,w,c", [(2, 5, 7, 64), (1, 22, 30, 1024), (2, 45, 60, 512), (1, 1, 3, 64), (2, 9, 11, 24)])
def test_bilinear2x_fused_with_producer_bn_relu_is_bit_identical(ops, cuda, n, h, w, c):
    """cvb_bn_relu_bilinear2x_fwd == cvb_bn_relu_apply + cvb_bilinear2x_fwd (the activation rounded to bf16 in between)."""
    y = _rand((n, h, w, c), cuda, 80).to(torch.bfloat16)
    scale = (torch.rand(c, device=cuda) + 0.5) * torch.where(torch.rand(c, device=cuda) < 0.2, -1.0, 1.0)
    shift = torch.randn(c, device=cuda) * 0.3
    a = torch.empty_like(y)
    ops.bn_relu_apply(y, scale, shift, a)
    want = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=cuda)
    ops.bilinear2x(a, want)
    got = torch.full_like(want, float("nan"))
    ops.bn_relu_bilinear2x(y, scale, shift, got)
    assert torch.equal(got, want)


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 24, 64), (1, 45, 61, 128), (3, 11, 15, 1024), (2, 9, 8, 16)])
@pytest.mark.parametrize("mode", ["unet", "segnet"])
def test_maxpool_bwd_fused_with_bn_reduce_is_bit_identical(ops, cuda, n, h, w, c, mode):
    """cvb_maxpool2x2_bwd_bn_reduce (cross-layer fusion, producer side) against the two kernels it replaces on the same
    buffers: dx bit-identical, reduction partials summing to the same (sum g, sum g*y) up to fp32 summation order.
    unet: accumulate into the skip gradient (the unfused reference recomputes the argmax from the activation, the fused
    kernel reads the forward's codes); segnet: plain scatter by the stored codes."""
    torch.manual_seed(33)
    y = _rand((n, h, w, c), cuda, 70).to(torch.bfloat16)
    scale = (torch.rand(c, device=cuda) + 0.5) * torch.where(torch.rand(c, device=cuda) < 0.2, -1.0, 1.0)  # some gamma < 0
    shift = torch.randn(c, device=cuda) * 0.3
    a = torch.empty_like(y)
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=cuda)
    code = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=cuda)
    ops.bn_relu_maxpool2x2(y, scale, shift, a, pooled, code)  # the forward that produced the activation and the codes
    dout = _rand((n, h // 2, w // 2, c), cuda, 71).to(torch.bfloat16)
    skip = _rand((n, h, w, c), cuda, 72).to(torch.bfloat16)
    rows = 3 * 37
    acc = mode == "unet"
    # reference: the unfused pair
    dx_ref = skip.clone() if acc else torch.full_like(skip, 7.0)
    ops.maxpool2x2_bwd(dout, dx_ref, code=None if acc else code, x=a if acc else None, accumulate=acc)
    parts_ref = torch.zeros(rows, 2, c, device=cuda)
    ops.bn_relu_bwd_reduce(dx_ref, y, scale, shift, parts_ref, rows)
    # fused
    dx = skip.clone() if acc else torch.full_like(skip, 7.0)
    parts = torch.full((rows, 2, c), float("nan"), device=cuda)
    ops.maxpool2x2_bwd_bn_reduce(dout, dx, y, scale, shift, parts, rows, code=code, accumulate=acc)
    assert torch.equal(dx, dx_ref)
    s, s_ref = parts.double().sum(0), parts_ref.double().sum(0)
    assert torch.isfinite(s).all()
    torch.testing.assert_close(s, s_ref, rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------------------------------- loss / metric
@pytest.mark.parametrize("hw", [(17, 23), (16, 24)])  # generic kernel / four-pixel vectorised 12-class kernel
@pytest.mark.parametrize("ignore", [-100, 11])
@pytest.mark.parametrize("labels", ["int64", "uint8"])
def test_softmax_ce(ops, cuda, ignore, hw, labels):
    torch.manual_seed(25)
    n, c, (h, w) = 3, 12, hw
    logits = F.relu(torch.randn(n, c, h, w) * 2).to(cuda).requires_grad_(True)
    target = torch.randint(0, c, (n, h, w)).to(cuda)
    if ignore == -100:
        target[0, :3] = -100 if labels == "int64" else 11  # ignored pixels with the default index too (int64 only)
    ref = F.cross_entropy(logits, target, ignore_index=ignore)
    (dref,) = torch.autograd.grad(ref, logits)
    tg = target if labels == "int64" else target.to(torch.uint8)
    dl = torch.empty_like(logits)
    loss = ops.softmax_ce_nchw(logits.detach(), tg, ignore, True, dl)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert rel_err(dl, dref) < 1e-5  # gradient scale = 1 / counted pixels, whatever ignore_index is
    # reduction='sum'
    ref_s = F.cross_entropy(logits, target, ignore_index=ignore, reduction="sum")
    (dref_s,) = torch.autograd.grad(ref_s, logits)
    loss_s = ops.softmax_ce_nchw(logits.detach(), tg, ignore, False, dl)
    assert abs(loss_s.item() - ref_s.item()) < 1e-5 * abs(ref_s.item())
    assert rel_err(dl, dref_s) < 1e-5
    # NHWC bf16 variant (model-internal logits, 64-channel padded gradient)
    lg = torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device=cuda)
    lg[..., :c] = to_nhwc_bf16(logits.detach())
    lq = to_nchw_f32(lg[..., :c]).requires_grad_(True)
    ref2 = F.cross_entropy(lq, target, ignore_index=ignore)
    (dref2,) = torch.autograd.grad(ref2, lq)
    dl2 = torch.full((n, h, w, 64), 1.0, dtype=torch.bfloat16, device=cuda)
    loss2 = ops.softmax_ce_nhwc(lg, c, tg, ignore, True, dl2)
    assert abs(loss2.item() - ref2.item()) < 1e-5 * max(1.0, abs(ref2.item()))
    assert rel_err(to_nchw_f32(dl2[..., :c]), dref2) < 4e-3
    assert torch.count_nonzero(dl2[..., c:]) == 0


def test_softmax_ce_flags_out_of_range_labels_and_all_ignored(ops, cuda):
    """torch raises a device-side assert on a label that is neither ignore_index nor in [0, C); across the C ABI the
    loss is poisoned with NaN instead (nothing is killed, nothing passes silently). All pixels ignored: 0/0 = NaN and
    a zero gradient, like torch."""
    torch.manual_seed(27)
    n, c, h, w = 2, 12, 16, 24
    logits = torch.randn(n, c, h, w, device=cuda)
    target = torch.randint(0, c, (n, h, w), device=cuda)
    dl = torch.empty_like(logits)
    assert torch.isfinite(ops.softmax_ce_nchw(logits, target, -100, True, dl))
    bad = target.clone()
    bad[1, 3, 5] = 255  # settings.IGNORE_LABEL (conf/settings.py:25) without telling the loss
    assert torch.isnan(ops.softmax_ce_nchw(logits, bad, -100, True, dl))
    assert torch.isfinite(ops.softmax_ce_nchw(logits, bad, 255, True, dl))
    allig = torch.full_like(target, 7)
    assert torch.isnan(ops.softmax_ce_nchw(logits, allig, 7, True, dl))
    assert torch.count_nonzero(dl) == 0


@pytest.mark.parametrize("hw", [(33, 47), (32, 44)])  # generic kernels / vectorised 12-class argmax kernel
def test_confusion_matrix_bit_exact(ops, cuda, hw):
    torch.manual_seed(26)
    n, c, (h, w) = 4, 12, hw
    logits = F.relu(torch.randn(n, c, h, w)).to(cuda)  # post-ReLU: argmax ties at 0 resolve to the first index
    gt = torch.randint(0, c, (n, h, w)).to(cuda)
    pred_ref = logits.argmax(1)
    cm_ref = torch.zeros(c, c, dtype=torch.int64, device=cuda)
    cm_ref.view(-1).index_add_(0, (gt * c + pred_ref).view(-1), torch.ones(gt.numel(), dtype=torch.int64, device=cuda))
    cm = torch.zeros(c, c, dtype=torch.int64, device=cuda)
    pred = torch.empty(n, h, w, dtype=torch.int64, device=cuda)
    ops.argmax_confusion_nchw(logits, gt, cm, pred)
    assert torch.equal(pred, pred_ref) and torch.equal(cm, cm_ref)
    cm.zero_()
    ops.confusion_matrix(pred_ref, gt, c, cm)
    assert torch.equal(cm, cm_ref)
    cm.zero_()
    lg = torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device=cuda)
    lg[..., :c] = to_nhwc_bf16(logits)
    pred2 = torch.empty_like(pred)
    ops.argmax_confusion_nhwc(lg, c, gt, cm, pred2)
    assert torch.equal(pred2, to_nchw_f32(lg[..., :c]).argmax(1))
    # uint8 ground truth (device-resident masks of the input stage): same counts
    cm.zero_()
    ops.argmax_confusion_nchw(logits, gt.to(torch.uint8), cm)
    assert torch.equal(cm, cm_ref)
    cm.zero_()
    ops.confusion_matrix(pred_ref.to(torch.uint8), gt.to(torch.uint8), c, cm)
    assert torch.equal(cm, cm_ref)


def test_pack_weights_batch_equals_single_layer_packers(ops, cuda):
    """One-launch packing of several layers == cvb_pack_weights_fprop / _dgrad per layer, bit for bit (incl. padding)."""
    shapes = [(64, 64), (12, 64), (128, 64), (256, 192)]  # (cout, cin); 12 -> padded to 64 output channels
    ws = [_rand((co, ci, 3, 3), cuda, 40 + i) for i, (co, ci) in enumerate(shapes)]
    entries, singles = [], []
    for w in ws:
        co_p, ci_p = ops.pad64(w.shape[0]), ops.pad64(w.shape[1])
        df = torch.full((co_p, 9 * ci_p), 7.0, dtype=torch.bfloat16, device=cuda)
        dd = torch.full((ci_p, 9 * co_p), 7.0, dtype=torch.bfloat16, device=cuda)
        entries.append((w, df, dd, co_p, ci_p))
        singles.append((ops.pack_weights_fprop(w, 9, co_p, ci_p), ops.pack_weights_dgrad(w, co_p, ci_p)))
    table = ops.pack_table(entries, cuda)
    ops.pack_weights_batch(table, max(e[3] for e in entries), max(e[4] for e in entries), 0)
    torch.cuda.synchronize()
    for (w, df, dd, _, _), (sf, sd) in zip(entries, singles):
        assert torch.equal(df, sf) and torch.equal(dd, sd), tuple(w.shape)


# ------------------------------------------------------------------------------------------- optimizer
def test_fused_adamw_matches_torch_adamw(cuda):
    """camvid_b200.optim.AdamW against torch.optim.AdamW (train.py:100) over several steps under OneCycleLR, which rewrites
    lr and beta1 every step (train.py:102-104,134): parameters and both moments within fp32 rounding of each other;
    state dicts interchangeable."""
    import camvid_b200  # noqa: F401
    from camvid_b200.optim import AdamW
    torch.manual_seed(61)
    shapes = [(64, 3, 3, 3), (64,), (128, 64, 3, 3), (12,), (5, 7), (100003,), (1,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device=cuda)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.AdamW(ref_p, lr=5e-4, weight_decay=1e-2)
    our = AdamW(our_p, lr=5e-4, weight_decay=1e-2)
    rs = torch.optim.lr_scheduler.OneCycleLR(ref, max_lr=5e-4, steps_per_epoch=4, epochs=2)
    os_ = torch.optim.lr_scheduler.OneCycleLR(our, max_lr=5e-4, steps_per_epoch=4, epochs=2)
    for it in range(6):
        flat = torch.randn(sum(p.numel() for p in ref_p), device=cuda)  # gradients = views of one flat buffer (engine)
        off = 0
        for a, b in zip(ref_p, our_p):
            gslice = flat[off:off + a.numel()].view_as(a)
            a.grad, b.grad = gslice.clone(), gslice
            off += a.numel()
        ref.step()
        our.step()
        rs.step()
        os_.step()
        for a, b in zip(ref_p, our_p):
            torch.testing.assert_close(b, a, rtol=2e-6, atol=1e-7)
    for a, b in zip(ref_p, our_p):
        torch.testing.assert_close(our.state[b]["exp_avg"], ref.state[a]["exp_avg"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(our.state[b]["exp_avg_sq"], ref.state[a]["exp_avg_sq"], rtol=1e-5, atol=1e-7)
        assert float(our.state[b]["step"]) == float(ref.state[a]["step"]) == 6.0
    fresh = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in our_p], lr=5e-4)
    fresh.load_state_dict(our.state_dict())  # same param_groups / state layout
    assert fresh.state_dict()["state"][0]["exp_avg"].shape == shapes[0]


# ------------------------------------------------------------------------------------------- input stage
def test_input_stage_bit_exact_with_reference_transforms(ops, cuda):
    """cvb_input_stage_u8 against the fixture produced by the reference's transforms.ToTensor + transforms.Normalize
    (transforms.py:485-538; tests/golden/make_golden.py) -- every byte value in every channel -- and against the oracle
    restatement on a CamVid-sized batch: bit-exact, image and mask."""
    import os
    import numpy as np
    from oracle import camvid_oracle as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_stage.npz"))
    mean, std = tuple(g["mean"]), tuple(g["std"])
    for tag in ("vec", "odd"):
        img, mask = torch.from_numpy(g[f"{tag}/img"]).to(cuda), torch.from_numpy(g[f"{tag}/mask"]).to(cuda)
        n, h, w, c = img.shape
        out = torch.full((n, c, h, w), float("nan"), device=cuda)
        m64 = torch.full((n, h, w), -1, dtype=torch.int64, device=cuda)
        ops.input_stage_u8(img, mean, std, out, mask, m64)
        assert torch.equal(out.cpu(), torch.from_numpy(g[f"{tag}/out"])), tag
        assert torch.equal(m64.cpu(), torch.from_numpy(g[f"{tag}/mask_out"])), tag
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (4, 360, 480, 3)).astype(np.uint8)
    mask = rng.integers(0, 12, (4, 360, 480)).astype(np.uint8)
    ref, ref_m = O.to_tensor_normalize(img, mask, mean, std)
    out = torch.empty(4, 3, 360, 480, device=cuda)
    m64 = torch.empty(4, 360, 480, dtype=torch.int64, device=cuda)
    ops.input_stage_u8(torch.from_numpy(img).to(cuda), mean, std, out, torch.from_numpy(mask).to(cuda), m64)
    assert torch.equal(out.cpu(), torch.from_numpy(ref)) and torch.equal(m64.cpu(), torch.from_numpy(ref_m))
    # mask only
    m64.fill_(-1)
    ops.input_stage_u8(None, (), (), None, torch.from_numpy(mask).to(cuda), m64)
    assert torch.equal(m64.cpu(), torch.from_numpy(ref_m))


@pytest.mark.parametrize("mask_dtype", [torch.int64, torch.uint8])
def test_device_prefetcher(cuda, mask_dtype):
    """camvid_b200.data.DevicePrefetcher over a loader of uint8 HWC batches (ragged last batch, pageable memory): every
    yielded batch equals the oracle's ToTensor + Normalize of the matching host batch, in order; h2d byte accounting;
    the reference's own fp32 / int64 batch format passes through unchanged."""
    import numpy as np
    import camvid_b200  # noqa: F401
    from camvid_b200 import data
    from oracle import camvid_oracle as O
    rng = np.random.default_rng(5)
    sizes = [4, 4, 4, 4, 3]
    batches = [(rng.integers(0, 256, (b, 36, 48, 3)).astype(np.uint8), rng.integers(0, 12, (b, 36, 48)).astype(np.uint8))
               for b in sizes]
    pf = data.DevicePrefetcher(batches, cuda, mask_dtype=mask_dtype)
    seen = 0
    for (img, mask), (x, m) in zip(batches, pf):
        ref, ref_m = O.to_tensor_normalize(img, mask, data.CAMVID_MEAN, data.CAMVID_STD)
        assert x.dtype == torch.float32 and m.dtype == mask_dtype and x.is_cuda and m.is_cuda
        torch.cuda.current_stream().synchronize()
        assert torch.equal(x.cpu(), torch.from_numpy(ref)), seen
        assert torch.equal(m.cpu().to(torch.int64), torch.from_numpy(ref_m)), seen
        seen += 1
    assert seen == len(sizes)
    assert pf.h2d_bytes == sum(i.size + m.size for i, m in batches)
    if mask_dtype == torch.int64:
        ref_batches = [(torch.randn(2, 3, 8, 12), torch.randint(0, 12, (2, 8, 12))) for _ in range(3)]
        for (img, mask), (x, m) in zip(ref_batches, data.DevicePrefetcher(ref_batches, cuda)):
            assert torch.equal(x.cpu(), img) and torch.equal(m.cpu(), mask)
