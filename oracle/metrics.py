"""Label-permutation-invariant parity metrics.  TEST INFRASTRUCTURE.

Restates the two assessment quantities the reference computes for a pair of
label volumes (src/iterseg/metrics.py:107 and :205-227):

* variation of information, as skimage.metrics.variation_of_information
  (conditional entropies in bits, log base 2 -- the more conservative choice
  for a "<= 0.01" gate), returned as the pair (H(seg|gt), H(gt|seg)); the gate
  uses their sum (background included, ignore_labels=()).
* IoU-matched true/false positives/negatives at a threshold (umetrix strict
  matching: one-to-one, IoU > threshold), from which matched-object F1 =
  2TP / (2TP + FP + FN).  Background (0) is not an object.
"""
import numpy as np
from scipy import sparse


def contingency(a, b):
    a = np.asarray(a).ravel().astype(np.int64)
    b = np.asarray(b).ravel().astype(np.int64)
    data = np.ones(a.size, dtype=np.int64)
    return sparse.coo_matrix((data, (a, b))).tocsr()


def variation_of_information(gt, seg):
    n = float(np.asarray(gt).size)
    pxy = contingency(gt, seg).astype(np.float64) / n
    px = np.asarray(pxy.sum(axis=1)).ravel()
    py = np.asarray(pxy.sum(axis=0)).ravel()
    coo = pxy.tocoo()
    v = coo.data
    # H(X|Y) = -sum pxy log2(pxy / py),  H(Y|X) = -sum pxy log2(pxy / px)
    hxgy = -np.sum(v * np.log2(v / py[coo.col]))
    hygx = -np.sum(v * np.log2(v / px[coo.row]))
    return float(hygx), float(hxgy)


def matched_counts(gt, seg, iou_threshold=0.5):
    c = contingency(gt, seg).tocoo()
    area_gt = np.bincount(np.asarray(gt).ravel().astype(np.int64))
    area_sg = np.bincount(np.asarray(seg).ravel().astype(np.int64))
    n_gt = int(np.count_nonzero(area_gt[1:]))
    n_sg = int(np.count_nonzero(area_sg[1:]))
    sel = (c.row > 0) & (c.col > 0)
    r, k, inter = c.row[sel], c.col[sel], c.data[sel].astype(np.float64)
    iou = inter / (area_gt[r] + area_sg[k] - inter)
    ok = iou > iou_threshold
    if iou_threshold >= 0.5:
        tp = int(ok.sum())          # IoU > 0.5 matches are necessarily one-to-one
    else:
        order = np.argsort(-iou[ok], kind='stable')
        used_r, used_k, tp = set(), set(), 0
        for i in order:
            a, b = r[ok][i], k[ok][i]
            if a not in used_r and b not in used_k:
                used_r.add(a); used_k.add(b); tp += 1
    return tp, n_sg - tp, n_gt - tp


def matched_f1(gt, seg, iou_threshold=0.5):
    tp, fp, fn = matched_counts(gt, seg, iou_threshold)
    d = 2 * tp + fp + fn
    return 1.0 if d == 0 else 2.0 * tp / d
