"""Shared helpers for the parity tests."""
import torch


def to_nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def to_nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (about half of the post-ReLU logits are exactly 0)."""
    a = a.double()
    b = b.double()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
