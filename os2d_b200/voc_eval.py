"""Detection evaluation on the device: drop-in for ``os2d.data.voc_eval.do_voc_evaluation``
(voc_eval.py:14-68), same arguments and result dict.

The reference walks images x labels in Python / numpy on the CPU (``boxes.append(boxes_one_image.cpu())``,
evaluate.py:118).  Here the detections stay where decode_pyramid left them: one kernel matches every detection to the
ground truth of its image and label (csrc/voc.cu), the greedy true / false positive flags, the per-label precision /
recall curves and both AP definitions are sorts, segmented scans and scatter-reductions over the whole dataset in
fp64; only the result arrays go back to the host.

Order of equal scores: the reference sorts with numpy's unstable ``argsort()[::-1]``; here ties keep the input order
(stable sorts).  With distinct scores inside a label the results equal the reference's to the last bit of the fp64
summation order (tests/test_gpu_voc.py compares with 1e-12); with tied scores AP can differ by the order of the ties.
"""
import ctypes

import numpy as np
import torch

from . import _cabi


def _flatten(predictions, gt_boxes, dev):
    det_b, det_l, det_s, det_i, gt_b, gt_l, gt_d, offs = [], [], [], [], [], [], [], [0]
    for i, (pred, gt) in enumerate(zip(predictions, gt_boxes)):
        pred = pred.resize(gt.image_size)                       # voc_eval.py:28-31
        n = pred.bbox_xyxy.shape[0]
        det_b.append(pred.bbox_xyxy.to(dev, torch.float32).reshape(-1, 4))
        det_l.append(pred.get_field("labels").to(dev).reshape(-1).to(torch.int64))
        det_s.append(pred.get_field("scores").to(dev, torch.float32).reshape(-1))
        det_i.append(torch.full((n,), i, dtype=torch.int32, device=dev))
        m = gt.bbox_xyxy.shape[0]
        gt_b.append(gt.bbox_xyxy.to(dev, torch.float32).reshape(-1, 4))
        gl = gt.get_field("labels").to(dev).reshape(-1).to(torch.int64)
        gt_l.append(gl)
        gt_d.append(gt.get_field("difficult").to(dev).reshape(-1).to(torch.bool) if gt.has_field("difficult")
                    else torch.zeros_like(gl, dtype=torch.bool))
        offs.append(offs[-1] + m)
    cat = lambda xs, dt: torch.cat(xs) if xs else torch.zeros(0, dtype=dt, device=dev)   # noqa: E731
    return (cat(det_b, torch.float32).contiguous(), cat(det_l, torch.int64), cat(det_s, torch.float32), cat(det_i, torch.int32),
            cat(gt_b, torch.float32).contiguous(), cat(gt_l, torch.int64), cat(gt_d, torch.bool),
            torch.tensor(offs, dtype=torch.int32, device=dev))


def _curves(label, score, match, n_pos, L):
    """Sorted (label asc, score desc) cumulative curves.  Returns order, counts [L], seg_start [L], prec, rec (fp64, in
    sorted order), tp_total [L]."""
    o1 = torch.argsort(score, descending=True, stable=True)
    order = o1[torch.argsort(label[o1], stable=True)]
    lab = label[order]
    m = match[order]
    counts = torch.bincount(lab, minlength=L)
    seg_start = torch.cumsum(counts, 0) - counts
    tp_c = torch.cumsum((m == 1).to(torch.float64), 0)
    fp_c = torch.cumsum((m == 0).to(torch.float64), 0)
    zero = torch.zeros(1, dtype=torch.float64, device=label.device)
    tp_base = torch.cat([zero, tp_c])[seg_start]           # cumulative value just before each segment
    fp_base = torch.cat([zero, fp_c])[seg_start]
    tp = tp_c - tp_base[lab]
    fp = fp_c - fp_base[lab]
    prec = tp / (fp + tp)                                   # 0/0 -> nan, as in the reference (voc_eval.py:163)
    rec = tp / n_pos[lab]
    seg_end = seg_start + counts
    tp_total = torch.cat([zero, tp_c])[seg_end] - tp_base
    return lab, counts, seg_start, prec, rec, tp_total


def _average_precision(lab, counts, seg_start, prec, rec, valid, L, use_07_metric):
    """voc_eval.py:171-230 for all labels at once.  valid [L]: label seen and n_pos > 0 (else AP = nan)."""
    dev = lab.device
    ap = torch.full((L,), float("nan"), dtype=torch.float64, device=dev)
    precz = torch.nan_to_num(prec, nan=0.0)
    rec = torch.where(valid[lab], rec, torch.zeros_like(rec))      # labels without positives: tp / 0, masked out below
    if use_07_metric:
        acc = torch.zeros(L, dtype=torch.float64, device=dev)
        for t in np.arange(0.0, 1.1, 0.1):
            sel = rec >= t
            p = torch.zeros(L, dtype=torch.float64, device=dev)
            p.scatter_reduce_(0, lab[sel], precz[sel], reduce="amax", include_self=True)
            acc = acc + p / 11
        ap[valid] = acc[valid]
        return ap
    n = lab.numel()
    # extended arrays with the sentinels of every label: [0, prec..., 0] and [0, rec..., 1]
    ext_n = n + 2 * L
    lab_all = torch.arange(L, device=dev)
    first = seg_start + 2 * lab_all                         # position of the leading sentinel of label l
    last = first + counts + 1
    pos = torch.arange(n, device=dev) + 2 * lab + 1
    mpre = torch.zeros(ext_n, dtype=torch.float64, device=dev)
    mrec = torch.zeros(ext_n, dtype=torch.float64, device=dev)
    elab = torch.empty(ext_n, dtype=torch.int64, device=dev)
    mpre[pos] = precz
    mrec[pos] = rec
    mrec[last] = 1.0
    elab[pos] = lab
    elab[first] = lab_all
    elab[last] = lab_all
    # per-label reverse running maximum: labels later in the array get a smaller offset, so they never leak backwards
    off = 2.0 * (L - elab).to(torch.float64)
    mpre = torch.flip(torch.cummax(torch.flip(mpre + off, [0]), 0).values, [0]) - off
    d = mrec[1:] - mrec[:-1]
    same = elab[1:] == elab[:-1]
    contrib = torch.where(same & (d != 0), d * mpre[1:], torch.zeros_like(d))
    csum = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), torch.cumsum(contrib, 0)])
    ap_all = csum[last] - csum[first]                       # contributions of pairs (k, k+1) with first <= k < last
    ap[valid] = ap_all[valid]
    return ap


@_cabi.on_device_of
def do_voc_evaluation(predictions, gt_boxes, iou_thresh=0.5, use_07_metric=False):
    """predictions / gt_boxes: lists of BoxList (one per image; fields "labels", "scores" / "labels", optional
    "difficult").  Returns the dict of the reference: ap_per_class, map, map_weighted, recall_per_class, recall, n_pos,
    prec, rec, ap_joint_classes."""
    assert len(gt_boxes) == len(predictions), "Length of gt and pred lists need to be same."
    if not torch.cuda.is_available():
        raise RuntimeError("os2d_b200 requires CUDA (no CPU path)")
    lib = _cabi.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    det_b, det_l, det_s, det_i, gt_b, gt_l, gt_d, offs = _flatten(predictions, gt_boxes, dev)
    n_det, n_gt = det_l.numel(), gt_l.numel()
    if n_det + n_gt == 0:
        raise ValueError("no detections and no ground truth")   # the reference fails on max() of an empty dict
    L = int(max(det_l.max().item() if n_det else -1, gt_l.max().item() if n_gt else -1)) + 1

    # ---- matching (kernel) + greedy flags ----
    gt_index = torch.empty(n_det, dtype=torch.int32, device=dev)
    det_l32, gt_l32 = det_l.to(torch.int32), gt_l.to(torch.int32)     # named: they must outlive the asynchronous launch
    _cabi.check(lib.os2d_voc_match(_cabi.ptr(det_b), _cabi.ptr(det_i), _cabi.ptr(det_l32), _cabi.ptr(gt_b), _cabi.ptr(gt_l32),
                                   _cabi.ptr(offs), n_det, ctypes.c_float(float(np.float32(iou_thresh))), _cabi.ptr(gt_index),
                                   _cabi.stream_ptr()), "os2d_voc_match")
    return _evaluate_from_matches(det_l, det_s, gt_index.to(torch.int64), gt_l, gt_d, L, use_07_metric)


def _evaluate_from_matches(det_l, det_s, gi, gt_l, gt_d, L, use_07_metric):
    """Everything after the matching kernel (device-agnostic tensor code): greedy flags, curves, AP, result dict."""
    dev = det_l.device
    n_det, n_gt = det_l.numel(), gt_l.numel()
    has = gi >= 0
    # rank of every detection in (label asc, score desc, input order): the first detection of a ground-truth box wins
    o1 = torch.argsort(det_s, descending=True, stable=True)
    order = o1[torch.argsort(det_l[o1], stable=True)]
    rank = torch.empty_like(order)
    rank[order] = torch.arange(n_det, device=dev)
    first_rank = torch.full((max(n_gt, 1),), n_det, dtype=torch.int64, device=dev)
    first_rank.scatter_reduce_(0, gi[has], rank[has], reduce="amin", include_self=True)
    match = torch.zeros(n_det, dtype=torch.int8, device=dev)
    gsafe = gi.clamp_min(0)
    if n_gt:
        is_first = has & (rank == first_rank[gsafe])
        match[is_first] = 1
        match[has & gt_d[gsafe]] = -1

    # ---- per-label statistics ----
    n_pos = torch.bincount(gt_l[~gt_d], minlength=L).to(torch.float64) if n_gt else torch.zeros(L, dtype=torch.float64, device=dev)
    seen = torch.zeros(L, dtype=torch.bool, device=dev)
    seen[det_l] = True
    seen[gt_l] = True
    valid = seen & (n_pos > 0)
    lab, counts, seg_start, prec, rec, tp_total = _curves(det_l, det_s, match, n_pos, L)
    ap = _average_precision(lab, counts, seg_start, prec, rec, valid, L, use_07_metric)
    recall_pc = torch.full((L,), float("nan"), dtype=torch.float64, device=dev)
    recall_pc[valid] = (tp_total / n_pos)[valid]            # last value of rec, 0.0 for a label without detections
    n_tot = n_pos[valid].sum()
    recall = float((n_pos[valid] * recall_pc[valid]).sum() / n_tot) if float(n_tot) > 0 else float("nan")

    # ---- all labels merged into one (ap_joint_classes) ----
    zlab = torch.zeros_like(det_l)
    npos1 = n_pos.sum().reshape(1)
    lab1, counts1, seg1, prec1, rec1, _ = _curves(zlab, det_s, match, npos1, 1)
    ap1 = _average_precision(lab1, counts1, seg1, prec1, rec1, npos1 > 0, 1, use_07_metric)

    # ---- results to the host, in the reference's types ----
    ap_np = ap.cpu().numpy()
    n_pos_np = n_pos.cpu().numpy()
    counts_np = counts.cpu().numpy()
    prec_np, rec_np = prec.cpu().numpy(), rec.cpu().numpy()
    seen_np = seen.cpu().numpy()
    starts = np.concatenate([[0], np.cumsum(counts_np)])
    prec_list = [prec_np[starts[l]:starts[l + 1]] if seen_np[l] else None for l in range(L)]
    rec_list = [rec_np[starts[l]:starts[l + 1]] if (seen_np[l] and n_pos_np[l] > 0) else None for l in range(L)]
    with np.errstate(invalid="ignore"):
        return {"ap_per_class": ap_np, "map": np.nanmean(ap_np), "map_weighted": np.nansum(ap_np * n_pos_np / n_pos_np.sum()),
                "recall_per_class": recall_pc.cpu().numpy(), "recall": recall, "n_pos": n_pos_np, "prec": prec_list,
                "rec": rec_list, "ap_joint_classes": float(ap1[0].item())}
