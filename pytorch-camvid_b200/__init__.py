"""camvid_b200: B200-native (sm_100a) implementation of the pytorch-camvid UNet / SegNet train + eval hot path.

The directory is named `pytorch-camvid_b200/` (the project name); it is imported as `camvid_b200` through the
one-line loader `camvid_b200.py` at the repository root.

Layout (mirrors the reference's own module tree for the hot path):
    models/unet.py, models/segnet.py   drop-in nn.Modules (same ctor signatures, parameter names, state_dict keys)
    utils.py                           get_model, mean_iou, intersect_and_union
    legacy/metrics.py                  Metrics
    nn.py                              CrossEntropyLoss drop-in on the fused loss kernel
    optim.py                           AdamW drop-in: one fused multi-tensor launch per step
    data.py                            host -> device input stage: pinned ring + ToTensor / Normalize on the GPU
    graph.py                           whole training step as one CUDA graph
    engine.py                          execution plans (buffers, kernel sequences) behind the modules
    parallel.py                        data-parallel gradient all-reduce (NCCL, side stream)
    ops.py / _lib.py                   operator layer over the C ABI (include/camvid_b200.h)
    csrc/                              CUDA kernels + C ABI  -> libcamvid_b200.so
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
