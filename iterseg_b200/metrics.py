"""Assessment metrics on the device: the two quantities of src/iterseg/metrics.py that gate the
end-to-end parity (variation of information, :107; IoU-matched TP/FP/FN, :205-227), same
definitions as the reference reaches through scikit-image / umetrix.

    variation_of_information(gt, seg) -> (H(seg|gt), H(gt|seg))      # bits, background included
    matched_counts(gt, seg, 0.5)      -> (tp, fp, fn)
    matched_f1(gt, seg, 0.5)          -> 2TP / (2TP + FP + FN)

Inputs: integer label volumes of equal shape (numpy or torch, host or device); one C-ABI call
(`isg_label_metrics`: pair sort + run-length contingency table) computes everything.
"""
import numpy as np
import torch

from . import _lib

__all__ = ['label_metrics', 'variation_of_information', 'matched_counts', 'matched_f1']


def _dev_u32(a, dev):
    if isinstance(a, torch.Tensor):
        t = a.to(dev)
        if t.dtype != torch.int32:
            t = t.to(torch.int64).to(torch.int32)
        return t.contiguous().view(-1)
    a = np.ascontiguousarray(a)
    if a.dtype not in (np.uint32, np.int32):
        a = a.astype(np.int64).astype(np.uint32)
    return torch.from_numpy(a.view(np.int32).reshape(-1)).to(dev)


def label_metrics(gt, seg, iou_threshold=0.5, max_label=None):
    """dict(vi_seg_given_gt, vi_gt_given_seg, tp, fp, fn, f1, n_seg, n_gt)."""
    _lib.require_device()
    lib = _lib.load()
    dev = gt.device if isinstance(gt, torch.Tensor) and gt.is_cuda else \
        torch.device('cuda', torch.cuda.current_device())
    g, s = _dev_u32(gt, dev), _dev_u32(seg, dev)
    if g.numel() != s.numel():
        raise ValueError('gt and seg must have the same number of voxels')
    n = g.numel()
    if max_label is None:
        max_label = int(max(int(g.view(torch.int32).max().item()), int(s.view(torch.int32).max().item()), 0))
    nb = lib.isg_metrics_workspace_bytes(n, max_label)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.isg_label_metrics(g.data_ptr(), s.data_ptr(), n, int(max_label), float(iou_threshold),
                                   out.data_ptr(), ws.data_ptr(), nb, _lib.stream_ptr())
    _lib.check(rc, 'isg_label_metrics')
    o = out.cpu().numpy()
    tp, n_sg, n_gt = int(round(o[2])), int(round(o[3])), int(round(o[4]))
    fp, fn = n_sg - tp, n_gt - tp
    d = 2 * tp + fp + fn
    return {'vi_seg_given_gt': float(o[0]), 'vi_gt_given_seg': float(o[1]), 'tp': tp, 'fp': fp, 'fn': fn,
            'f1': 1.0 if d == 0 else 2.0 * tp / d, 'n_seg': n_sg, 'n_gt': n_gt}


def variation_of_information(gt, seg):
    m = label_metrics(gt, seg)
    return m['vi_seg_given_gt'], m['vi_gt_given_seg']


def matched_counts(gt, seg, iou_threshold=0.5):
    m = label_metrics(gt, seg, iou_threshold)
    return m['tp'], m['fp'], m['fn']


def matched_f1(gt, seg, iou_threshold=0.5):
    return label_metrics(gt, seg, iou_threshold)['f1']
