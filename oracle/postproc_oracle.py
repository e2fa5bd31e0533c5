"""CPU restatement (numpy + C greedy NMS) of box decoding and per-class NMS across pyramid levels.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
Follows os2d/modeling/box_coder.py:448-536 (decode_pyramid), :319-330 (decode_single wrapper), :424-437
(_nms_box_lists); os2d/structures/bounding_box.py:344-387 (chunked NMS loop), :138-163 (resize), :261-281 (clip /
empty mask); torchvision models/detection/_utils.py BoxCoder.decode_single and ops/boxes.py clip_boxes_to_image
(third-party, not vendored - restated from the installed torchvision 0.26.0).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NMS_MAX_BATCH = 10000
BBOX_XFORM_CLIP = math.log(1000.0 / 16)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libos2d_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.os2d_oracle_nms.restype = ctypes.c_int64
        _LIB.os2d_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p]
    return _LIB


def greedy_nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms semantics on fp32 numpy arrays: returns kept indices, score-descending
    (stable for ties: the lower index first)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.ascontiguousarray(np.argsort(-scores, kind="stable").astype(np.int64))
    keep = np.empty(n, dtype=np.int64)
    k = _lib().os2d_oracle_nms(boxes.ctypes.data, order.ctypes.data, n, float(iou_threshold), keep.ctypes.data)
    return keep[:k].copy()


def chunked_nms(boxes, scores, iou_threshold, score_threshold=float("-inf")):
    """bounding_box.py:344-374: survivors (in order) split in chunks of 10000, greedy NMS per chunk, concatenated;
    repeated until there was at most one chunk or nothing was removed."""
    ids = np.nonzero(scores > np.float32(score_threshold))[0].astype(np.int64)
    while True:
        n = ids.shape[0]
        n_batches = int(math.ceil(n / float(NMS_MAX_BATCH)))
        survived = []
        for s in range(0, n, NMS_MAX_BATCH):
            b = ids[s:s + NMS_MAX_BATCH]
            survived.append(b[greedy_nms(boxes[b], scores[b], iou_threshold)])
        ids = np.concatenate(survived) if survived else np.zeros(0, dtype=np.int64)
        if n_batches <= 1 or ids.shape[0] == n:
            return ids


def anchors_xyxy(fm_w, fm_h, stride_w=16, stride_h=16, box_w=240, box_h=240):
    """box_coder.py:16-60: centres (i + 0.5) * stride, row-major over the feature map, xyxy = c -/+ size/2."""
    cx = (np.arange(fm_w, dtype=np.float32) + np.float32(0.5)) * np.float32(stride_w)
    cy = (np.arange(fm_h, dtype=np.float32) + np.float32(0.5)) * np.float32(stride_h)
    cx = np.tile(cx[None, :], (fm_h, 1)).reshape(-1)
    cy = np.tile(cy[:, None], (1, fm_w)).reshape(-1)
    hw, hh = np.float32(box_w / 2), np.float32(box_h / 2)
    return np.stack([cx - hw, cy - hh, cx + hw, cy + hh], axis=1).astype(np.float32)


def decode_boxes(loc, anchors):
    """torchvision BoxCoder.decode_single with weights (10,10,5,5).  loc [N,4], anchors [N,4] -> [N,4]."""
    f = np.float32
    loc = loc.astype(f)
    widths = anchors[:, 2] - anchors[:, 0]
    heights = anchors[:, 3] - anchors[:, 1]
    ctr_x = anchors[:, 0] + f(0.5) * widths
    ctr_y = anchors[:, 1] + f(0.5) * heights
    dx = loc[:, 0] / f(10)
    dy = loc[:, 1] / f(10)
    dw = np.minimum(loc[:, 2] / f(5), f(BBOX_XFORM_CLIP))
    dh = np.minimum(loc[:, 3] / f(5), f(BBOX_XFORM_CLIP))
    pcx = dx * widths + ctr_x
    pcy = dy * heights + ctr_y
    pw = np.exp(dw).astype(f) * widths
    ph = np.exp(dh).astype(f) * heights
    cw = f(0.5) * pw
    ch = f(0.5) * ph
    return np.stack([pcx - cw, pcy - ch, pcx + cw, pcy + ch], axis=1).astype(f)


def _resize(arr_xyxy, src_wh, dst_wh):
    """BoxList.resize (bounding_box.py:138-163) on an [n, 4k] array of x,y,x,y,... coordinates."""
    rw = float(dst_wh[0]) / src_wh[0]
    rh = float(dst_wh[1]) / src_wh[1]
    if rw == rh:
        return (arr_xyxy * np.float32(rw)).astype(np.float32)
    out = arr_xyxy.copy()
    out[:, 0::2] = arr_xyxy[:, 0::2] * np.float32(rw)
    out[:, 1::2] = arr_xyxy[:, 1::2] * np.float32(rh)
    return out.astype(np.float32)


def decode_pyramid(loc_pyr, cls_pyr, img_sizes, fm_sizes, class_ids, nms_score_threshold=0.0, nms_iou_threshold=0.3,
                   target_size=None, corners_pyr=None, nms_across_classes=False, box=240, stride=16):
    """box_coder.py:448-536.  loc_pyr[l] [C,4,N_l], cls_pyr[l] [C,N_l], img_sizes[l] = (w,h) of the level image,
    fm_sizes[l] = (w,h) of its feature map, target_size = (w,h) the inverse transforms map every level to (None:
    no inverse transform).  Returns dict of numpy arrays: boxes, scores, labels, default_boxes[, transform_corners]."""
    out_b, out_s, out_l, out_d, out_c = [], [], [], [], []
    for real_label in set(class_ids):
        bl, sl, dl, cl = [], [], [], []
        for i_label, cid in enumerate(class_ids):
            if cid != real_label:
                continue
            for lvl, (loc, cls) in enumerate(zip(loc_pyr, cls_pyr)):
                iw, ih = img_sizes[lvl]
                anc = anchors_xyxy(fm_sizes[lvl][0], fm_sizes[lvl][1], stride, stride, box, box)
                b = decode_boxes(np.ascontiguousarray(loc[i_label].T), anc)
                b[:, 0::2] = np.clip(b[:, 0::2], 0, iw)
                b[:, 1::2] = np.clip(b[:, 1::2], 0, ih)
                s = cls[i_label].astype(np.float32)
                bad = (b[:, 3] <= b[:, 1]) | (b[:, 2] <= b[:, 0])
                m = (s > np.float32(nms_score_threshold)) & ~bad
                if not m.any():
                    continue
                bm, am = b[m], anc[m]
                cm = np.ascontiguousarray(corners_pyr[lvl][i_label].T)[m] if corners_pyr is not None else None
                if target_size is not None:
                    bm = _resize(bm, (iw, ih), target_size)
                    am = _resize(am, (iw, ih), target_size)
                    if cm is not None:
                        cm = _resize(cm, (iw, ih), target_size)
                bl.append(bm); sl.append(s[m]); dl.append(am)
                if cm is not None:
                    cl.append(cm)
        if not bl:
            continue
        b = np.concatenate(bl); s = np.concatenate(sl); d = np.concatenate(dl)
        keep = chunked_nms(b, s, nms_iou_threshold)
        keep = keep[np.argsort(-s[keep], kind="stable")]
        out_b.append(b[keep]); out_s.append(s[keep]); out_d.append(d[keep])
        out_l.append(np.full(keep.shape[0], int(real_label), dtype=np.int64))
        if cl:
            out_c.append(np.concatenate(cl)[keep])
    res = {"boxes": np.concatenate(out_b) if out_b else np.zeros((0, 4), np.float32),
           "scores": np.concatenate(out_s) if out_s else np.zeros((0,), np.float32),
           "labels": np.concatenate(out_l) if out_l else np.zeros((0,), np.int64),
           "default_boxes": np.concatenate(out_d) if out_d else np.zeros((0, 4), np.float32)}
    if corners_pyr is not None:
        res["transform_corners"] = np.concatenate(out_c) if out_c else np.zeros((0, 8), np.float32)
    if nms_across_classes and res["boxes"].shape[0] > 0:
        keep = chunked_nms(res["boxes"], res["scores"], nms_iou_threshold)
        keep = keep[np.argsort(-res["scores"][keep], kind="stable")]
        res = {k: v[keep] for k, v in res.items()}
    return res
