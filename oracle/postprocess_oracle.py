"""CPU oracle for the post-processing half of the hot path (TEST INFRASTRUCTURE — not product).

NumPy restatement of ``yolov3/inference.py`` of nrsyed/pytorch-yolov3: the thresholding /
scaling / truncation of ``inference()`` (:338-366), ``cxywh_to_tlbr`` (:269-283) and the greedy
NMS (``_non_max_suppression`` :161-217, ``non_max_suppression`` :220-266).  Integer work is
int64, the IoU ratio is a float64 divide, exactly as NumPy evaluates the reference's
expressions.  Pinned against the live reference by tests/golden/make_golden.py.
"""
import numpy as np


def cxywh_to_tlbr(b):
    """inference.py:269-283 — tl = c - wh//2, br = c + wh//2; extra columns pass through."""
    out = np.copy(b)
    half = b[:, 2:4] // 2
    out[:, 0:2] = b[:, 0:2] - half
    out[:, 2:4] = b[:, 0:2] + half
    return out


def nms_single(tlbr, prob, iou_thresh=0.3):
    """inference.py:161-217.  Vectorised over the 'remaining' set but arithmetically identical:
    visit by descending prob (np.argsort(prob)[::-1]); keep the head; drop every remaining box
    whose inter/union (float64) is > iou_thresh."""
    tlbr = np.asarray(tlbr)
    x1, y1, x2, y2 = (tlbr[:, k] for k in range(4))
    area = ((x2 - x1) + 1) * ((y2 - y1) + 1)
    order = np.argsort(prob)[::-1]
    keep = []
    while order.size:
        cur, rest = order[0], order[1:]
        keep.append(int(cur))
        iw = np.maximum(0, (np.minimum(x2[cur], x2[rest]) - np.maximum(x1[cur], x1[rest])) + 1)
        ih = np.maximum(0, (np.minimum(y2[cur], y2[rest]) - np.maximum(y1[cur], y1[rest])) + 1)
        inter = iw * ih
        union = area[cur] + area[rest] - inter
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / union
        order = rest[~(iou > iou_thresh)]
    return keep


def nms(tlbr, prob, class_idx=None, iou_thresh=0.3):
    """inference.py:220-266 — per class in ``set(class_idx)`` iteration order (hash order, F6),
    indices mapped back to the caller's arrays; class-agnostic when class_idx is None."""
    if class_idx is None:
        return nms_single(tlbr, prob, iou_thresh)
    keep = []
    for c in set(class_idx):
        members = np.where(class_idx == c)[0]
        kept = nms_single(tlbr[members], prob[members], iou_thresh)
        keep.extend(members[kept].tolist())
    return keep


def postprocess(bbox_xywh, class_prob, class_idx, orig_shapes, prob_thresh=0.05, nms_iou_thresh=0.3):
    """inference.py:342-366 on host arrays: mask prob >= thresh, scale x,w by W and y,h by H in
    float32, truncate to int64, tl/br, per-class NMS.  Returns the reference's result list."""
    results = []
    mask = class_prob >= prob_thresh
    for i in range(bbox_xywh.shape[0]):
        box = bbox_xywh[i, mask[i], :].copy()
        prob = class_prob[i, mask[i]]
        idx = class_idx[i, mask[i]]
        box[:, [0, 2]] *= orig_shapes[i][1]
        box[:, [1, 3]] *= orig_shapes[i][0]
        tlbr = cxywh_to_tlbr(box.astype(np.int64))
        keep = nms(tlbr, prob, class_idx=idx, iou_thresh=nms_iou_thresh)
        results.append([tlbr[keep, :], prob[keep], idx[keep]])
    return results


def preprocess(images):
    """inference.py:332-333 — stack, BGR->RGB, HWC->CHW, float32, /255."""
    return np.transpose(np.flip(np.stack(images), 3), (0, 3, 1, 2)).astype(np.float32) / 255.0
