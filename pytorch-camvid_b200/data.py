"""Host -> device input stage of the training / eval loop.

The reference moves every batch with `images = images.cuda(); masks = masks.cuda()` (train.py:126-127, eval.py:52-53):
pageable fp32 NCHW images + int64 masks, 55 MB per 16 x 3 x 360 x 480 batch, synchronous, after the CPU workers ran
transforms.ToTensor and transforms.Normalize (transforms.py:485-538) on every image. `DevicePrefetcher` replaces
those two lines:

  * it owns a ring of PINNED staging buffers and device buffers and copies batch i+1 on a side stream while step i
    computes; the consumer's stream only waits on an event;
  * when the loader yields what cv2 and the dataset produce -- uint8 HWC images and uint8 masks, i.e. the reference's
    transform pipeline with its last two entries (ToTensor, Normalize) removed -- only those 11 MB cross PCIe and
    ToTensor + Normalize (+ the mask's .long()) run on the GPU in one kernel (cvb_input_stage_u8), bit-exact with the
    reference transforms;
  * batches in the reference's own format (fp32 NCHW + int64) are accepted too and just take the pinned ring.

    loader = DataLoader(dataset_without_totensor_normalize, batch_size=16, ...)
    for images, masks in DevicePrefetcher(loader, "cuda", mean=settings.MEAN, std=settings.STD):
        loss = loss_fn(net(images), masks)        # images fp32 [B,3,H,W], masks int64 (or uint8) [B,H,W], on the GPU

The yielded tensors belong to the ring: they are valid until the batch after the next one is requested (depth = 2),
like any CUDA prefetcher; clone them to keep them longer.
"""
import collections

import numpy as np
import torch

from . import ops

CAMVID_MEAN = (0.42019099703461577, 0.41323568513979647, 0.4010048431259079)  # conf/settings.py:8 (BGR)
CAMVID_STD = (0.30598050258519743, 0.3089986932156864, 0.3054061869915674)    # conf/settings.py:9


def _as_tensor(a):
    if torch.is_tensor(a):
        return a
    if isinstance(a, (list, tuple)):  # a list of per-image arrays: the default collate of numpy samples stacks them
        return torch.from_numpy(np.stack([np.asarray(x) for x in a]))
    return torch.from_numpy(np.ascontiguousarray(a))


class _Slot:
    def __init__(self):
        self.shape = None
        self.ready = torch.cuda.Event()  # the batch staged in this slot is complete on the device
        self.free = None                 # recorded on the consumer's stream once the slot's batch has been used


class DevicePrefetcher:
    """Iterates `loader`, yielding (images fp32 [B,C,H,W], masks [B,H,W]) on `device`.

    mean / std: per-channel normalisation of the uint8 path (defaults: CamVid BGR, conf/settings.py:8-9).
    mask_dtype: torch.int64 (default: what the reference's loss and metrics take) or torch.uint8 (kept as uploaded;
        camvid_b200.nn.CrossEntropyLoss, Metrics.add_logits and ops.argmax_confusion_nchw read uint8 labels directly,
        saving the widening pass).
    depth: ring size (>= 2)."""

    def __init__(self, loader, device="cuda", mean=CAMVID_MEAN, std=CAMVID_STD, mask_dtype=torch.int64, depth=2):
        self.loader, self.device = loader, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("camvid_b200.data.DevicePrefetcher stages batches onto a CUDA device; there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if mask_dtype not in (torch.int64, torch.uint8):
            raise ValueError("mask_dtype must be torch.int64 or torch.uint8")
        if depth < 2:
            raise ValueError("depth must be >= 2 (one batch in use, one in flight)")
        self.mean, self.std, self.mask_dtype, self.depth = tuple(mean), tuple(std), mask_dtype, depth
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = [_Slot() for _ in range(depth)]
        self.h2d_bytes = 0  # bytes copied host -> device so far (bench.py reports the per-step figure)

    def __len__(self):
        return len(self.loader)

    # ---- one batch into one slot, asynchronously on the side stream
    def _alloc(self, slot, key, img, mask):
        slot.shape = key
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        dev = lambda t, dt=None: torch.empty(t.shape, dtype=dt or t.dtype, device=self.device)
        slot.pin_img, slot.pin_mask = pin(img), pin(mask)
        slot.dev_img, slot.dev_mask = dev(img), dev(mask)
        if img.dtype == torch.uint8:
            n, h, w, c = img.shape
            slot.out_img = torch.empty(n, c, h, w, dtype=torch.float32, device=self.device)
        else:
            slot.out_img = slot.dev_img
        widen = mask.dtype == torch.uint8 and self.mask_dtype == torch.int64
        slot.out_mask = dev(mask, torch.int64) if widen else slot.dev_mask

    def _stage(self, slot, batch):
        img, mask = _as_tensor(batch[0]), _as_tensor(batch[1])
        u8 = img.dtype == torch.uint8
        if u8:
            if img.dim() != 4 or img.shape[3] > 4 or len(self.mean) != img.shape[3]:
                raise RuntimeError(f"uint8 image batches must be [B,H,W,C] with C == len(mean), got {tuple(img.shape)}")
            if mask.dtype != torch.uint8:
                mask = mask.to(torch.uint8) if mask.dtype in (torch.int64, torch.int32, torch.int16) and \
                    int(mask.max()) < 256 and int(mask.min()) >= 0 else mask
        elif not (img.dtype == torch.float32 and img.dim() == 4):
            raise RuntimeError(f"image batches must be uint8 [B,H,W,C] or fp32 [B,C,H,W], got {img.dtype} {tuple(img.shape)}")
        if mask.dtype not in (torch.uint8, torch.int64):
            raise RuntimeError(f"mask batches must be uint8 or int64 [B,H,W], got {mask.dtype}")
        if mask.dtype == torch.int64 and self.mask_dtype == torch.uint8:
            raise RuntimeError("mask_dtype=torch.uint8 needs uint8 masks from the loader")
        key = (tuple(img.shape), img.dtype, tuple(mask.shape), mask.dtype)
        if slot.shape != key:
            slot.ready.synchronize()
            self._alloc(slot, key, img, mask)
        slot.ready.synchronize()  # the previous copy out of this slot's pinned buffers has finished (long ago)
        # pageable -> pinned on this thread (a loader with pin_memory=True hands over pinned tensors: used in place)
        src_img = img if img.is_pinned() else slot.pin_img.copy_(img)
        src_mask = mask if mask.is_pinned() else slot.pin_mask.copy_(mask)
        if slot.free is not None:
            self.stream.wait_event(slot.free)  # the step that last read this slot's device buffers is done
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            slot.dev_img.copy_(src_img, non_blocking=True)
            slot.dev_mask.copy_(src_mask, non_blocking=True)
            self.h2d_bytes += img.numel() * img.element_size() + mask.numel() * mask.element_size()
            if u8:
                ops.input_stage_u8(slot.dev_img, self.mean, self.std, slot.out_img, slot.dev_mask,
                                   slot.out_mask if slot.out_mask is not slot.dev_mask else None)
            elif slot.out_mask is not slot.dev_mask:
                ops.input_stage_u8(None, (), (), None, slot.dev_mask, slot.out_mask)
            slot.ready.record(self.stream)
        slot.keep = (src_img, src_mask)  # pinned sources must outlive the asynchronous copy

    def __iter__(self):
        it = iter(self.loader)
        free = collections.deque(self.slots)
        staged = collections.deque()
        in_use = None

        def fill():
            while free:
                try:
                    batch = next(it)
                except StopIteration:
                    return
                slot = free.popleft()
                self._stage(slot, batch)
                staged.append(slot)

        fill()
        while staged:
            slot = staged.popleft()
            cur = torch.cuda.current_stream(self.device)
            if in_use is not None:  # everything enqueued so far has used the previous batch: its slot can be refilled
                in_use.free = torch.cuda.Event()
                in_use.free.record(cur)
                free.append(in_use)
            cur.wait_event(slot.ready)
            in_use = slot
            fill()  # batch i+1 (and beyond, ring permitting) starts copying before step i is even enqueued
            yield slot.out_img, slot.out_mask
