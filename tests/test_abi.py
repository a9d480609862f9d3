"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly the symbols the
header declares; the ctypes table covers all of them; the host-side mirrors keep the reference's interface."""
import os
import re
import subprocess

import pytest
import torch

import camvid_b200
from camvid_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "camvid_b200.h")).read()
    return sorted(set(re.findall(r"^CVB_API\s+[\w\s\*]+?\b(cvb_\w+)\s*\(", src, flags=re.M)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.cvb_abi_version() == _lib.ABI_VERSION == 4
    declared = _header_symbols()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(cvb_\w+)", out)))
    assert exported == declared
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name)


def test_library_has_no_torch_or_libcuda_link_dependency():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out and "libcuda.so" not in out


def test_error_reporting_without_compute():
    lib = _lib.load()
    v = _lib.View(None, 1, 1, 1, 8, 8, 8, 8)
    rc = lib.cvb_zero_view(v, None)
    assert rc == -1 and b"null pointer" in lib.cvb_last_error()
    with pytest.raises(RuntimeError, match="null pointer"):
        _lib.check(rc, "zero_view")


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in out.stdout and "UTMALDG" in out.stdout and "LDTM" in out.stdout


def test_get_model_interface_and_state_dict_keys():
    from camvid_b200.utils import get_model
    with pytest.raises(ValueError, match="network type does not supported"):
        get_model("deeplab", 3, 12)
    u, s = get_model("unet", 3, 12), get_model("segnet", 3, 12)
    assert sum(p.numel() for p in u.parameters()) == 34533924  # README.md:39
    assert sum(p.numel() for p in s.parameters()) == 29449956  # README.md:40
    uk, sk = list(u.state_dict()), list(s.state_dict())
    assert uk[0] == "down1.0.conv.0.weight" and uk[-1] == "output.conv.1.num_batches_tracked" and len(uk) == 161
    assert sk[0] == "encoder1.0.conv.weight" and sk[-1] == "decoder1.1.bn.num_batches_tracked" and len(sk) == 182
    names = [n for n, _ in u.named_parameters()]
    assert names[-1] == "output.conv.1.bias" and len(names) == 92  # utils.py:15-31 reads the last weight/bias
    assert [n for n, _ in s.named_parameters()][-1] == "decoder1.1.bn.bias"


def test_cpu_input_fails_loudly():
    from camvid_b200.utils import get_model
    net = get_model("unet", 3, 12)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(torch.zeros(1, 3, 32, 32))


def test_loss_module_rejects_unsupported_options():
    from camvid_b200.nn import CrossEntropyLoss
    CrossEntropyLoss()
    CrossEntropyLoss(ignore_index=11, reduction="sum")
    with pytest.raises(ValueError):
        CrossEntropyLoss(weight=torch.ones(12))
    with pytest.raises(ValueError):
        CrossEntropyLoss(label_smoothing=0.1)


def test_wgrad_workspace_planner_is_host_only_and_consistent():
    """cvb_conv3x3_wgrad_workspace_bytes runs the whole work-partition planner (stream-K ranges, lockstep waves) on the
    host: it must answer without a GPU, and the workspace is a whole number of [taps][cin_pad][cout_pad] fp32 slots, at most
    one per CTA (148 SMs assumed when no device is visible)."""
    import ctypes
    from camvid_b200 import _lib
    lib = _lib.load()
    for n, h, w, cin, cout in [(16, 360, 480, 64, 64), (16, 180, 240, 256, 128), (16, 90, 120, 512, 256),
                               (16, 45, 60, 1024, 512), (16, 22, 30, 1024, 1024), (1, 11, 15, 512, 512), (2, 8, 8, 32, 64), (2, 40, 72, 64, 16)]:
        x = _lib.View(ctypes.c_void_p(256), n, h, w, cin, h * w * cin, w * cin, cin)
        dy = _lib.View(ctypes.c_void_p(256), n, h, w, cout, h * w * cout, w * cout, cout)
        taps = 9 if cin >= 64 else 1
        nbytes = lib.cvb_conv3x3_wgrad_workspace_bytes(x, dy, taps)
        slot = taps * ((cin + 63) // 64 * 64) * ((cout + 63) // 64 * 64) * 4
        assert nbytes > 0 and nbytes % slot == 0 and 1 <= nbytes // slot <= 148, (n, h, w, cin, cout, nbytes)
