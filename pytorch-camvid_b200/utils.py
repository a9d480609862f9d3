"""Host-side mirror of the reference's utils.py for the hot path: get_model, intersect_and_union, mean_iou.

Same names, argument meaning, return values and error behaviour as utils.py:147-228 of the reference; the pixel
counting runs in the confusion-matrix kernel (C ABI cvb_confusion_matrix) instead of three np.histogram calls per
image, and accepts CUDA tensors directly (the reference copies predictions and masks to the host first, train.py:192).
"""
import numpy as np
import torch

from . import ops


def get_model(model_name, input_channels, class_num):
    """utils.py:147-160 -- lazy imports, ValueError for an unknown name."""
    if model_name == 'unet':
        from .models.unet import UNet
        return UNet(input_channels, class_num)
    if model_name == 'segnet':
        from .models.segnet import SegNet
        return SegNet(input_channels, class_num)
    raise ValueError('network type does not supported')


def _to_cuda_i64(a):
    t = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("camvid_b200 metrics run on CUDA only; no CPU path")
        t = t.cuda(non_blocking=True)
    return t.to(torch.int64).contiguous()


def _stack(maps):
    """list of per-image maps (ndarray / tensor) or one batched tensor -> one flat CUDA int64 tensor."""
    if torch.is_tensor(maps):
        return _to_cuda_i64(maps).view(-1)
    if isinstance(maps, np.ndarray):
        return _to_cuda_i64(maps).view(-1)
    return torch.cat([_to_cuda_i64(m).view(-1) for m in maps])


def _areas(pred, label, num_classes, ignore_index):
    """Pixel areas exactly as utils.py:178-188 computes them with np.histogram(bins=arange(C+1)):
    pixels whose label equals ignore_index are dropped; each histogram drops values outside [0, C] on its own, and
    the last bin is closed, so the value C is counted in class C-1."""
    C = num_classes
    ext = C + 2  # classes 0..C-1, the value C (np.histogram's closed last bin), everything else
    if label.device != pred.device:
        label = label.to(pred.device)
    with ops.on_device(pred):
        cm = torch.zeros(ext, ext, dtype=torch.int64, device=pred.device)
        ops.confusion_matrix(pred, label, ext, cm, ignore_label=ignore_index, clamp_oob=True)
    cm = cm.cpu().numpy().astype(np.float64)
    inter = np.diag(cm)[:C].copy()
    inter[C - 1] += cm[C, C]
    area_pred = cm.sum(axis=0)[:C].copy()
    area_pred[C - 1] += cm[:, C].sum()
    area_label = cm.sum(axis=1)[:C].copy()
    area_label[C - 1] += cm[C, :].sum()
    return inter, area_pred + area_label - inter, area_pred, area_label


def intersect_and_union(pred_label, label, num_classes, ignore_index):
    """utils.py:162-190: returns (area_intersect, area_union, area_pred_label, area_label), each of shape (C,)."""
    p, g = _to_cuda_i64(pred_label).view(-1), _to_cuda_i64(label).view(-1)
    inter, union, ap, al = _areas(p, g, num_classes, ignore_index)
    as_int = lambda a: a.astype(np.int64)  # np.histogram returns integer counts
    return as_int(inter), as_int(union), as_int(ap), as_int(al)


def mean_iou(results, gt_seg_maps, num_classes, ignore_index, nan_to_num=None):
    """utils.py:193-228: (all_acc, acc[C], iou[C]) with areas summed over the images; 0/0 stays NaN unless
    nan_to_num is given."""
    num_imgs = len(results)
    assert len(gt_seg_maps) == num_imgs
    p, g = _stack(results), _stack(gt_seg_maps)
    if p.numel() != g.numel():
        raise ValueError("results and gt_seg_maps hold a different number of pixels")
    inter, union, _, area_label = _areas(p, g, num_classes, ignore_index)
    with np.errstate(divide='ignore', invalid='ignore'):
        all_acc = inter.sum() / area_label.sum()
        acc = inter / area_label
        iou = inter / union
    if nan_to_num is not None:
        return all_acc, np.nan_to_num(acc, nan=nan_to_num), np.nan_to_num(iou, nan=nan_to_num)
    return all_acc, acc, iou
