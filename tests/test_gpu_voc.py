"""GPU: on-device detection evaluation (os2d_b200.voc_eval.do_voc_evaluation) against the
golden outputs of the reference's do_voc_evaluation and against the numpy oracle on a larger seeded dataset."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, voc_inputs

pytestmark = pytest.mark.gpu


def _boxlists(data, device):
    from os2d_b200.structures import BoxList, FeatureMapSize
    preds, gts = [], []
    for (pb, pl, ps, psize, gt, gl, gd, gsize) in data:
        b = BoxList(pb.to(device), FeatureMapSize(w=psize[0], h=psize[1]))
        b.add_field("labels", pl.to(device))
        b.add_field("scores", ps.to(device))
        preds.append(b)
        t = BoxList(gt, FeatureMapSize(w=gsize[0], h=gsize[1]))
        t.add_field("labels", gl)
        t.add_field("difficult", gd)
        gts.append(t)
    return preds, gts


def _check(r, ref, exact=True):
    np.testing.assert_array_equal(r["n_pos"], ref["n_pos"])
    tol = 1e-12 if exact else 0.1       # 16 score levels: heavy ties, AP depends on their (unspecified) order
    np.testing.assert_allclose(r["ap_per_class"], ref["ap_per_class"], rtol=0, atol=tol, equal_nan=True)
    np.testing.assert_allclose(r["recall_per_class"], ref["recall_per_class"], rtol=0, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose([r["map"], r["map_weighted"], r["ap_joint_classes"]],
                               [ref["map"], ref["map_weighted"], ref["ap_joint_classes"]], rtol=0, atol=tol)
    np.testing.assert_allclose(r["recall"], ref["recall"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_voc_eval_matches_reference_golden(device):
    """Detections handed over on the GPU (as decode_pyramid leaves them) or on the CPU (as evaluate.py:118 stores them)."""
    from os2d_b200.voc_eval import do_voc_evaluation
    gold = np.load(os.path.join(GOLDEN, "voc_eval.npz"))
    preds, gts = _boxlists(voc_inputs(91), device)
    for thr in (0.5, 0.3):
        for m07 in (False, True):
            r = do_voc_evaluation(preds, gts, iou_thresh=thr, use_07_metric=m07)
            key = "distinct_thr{}_{}".format(int(thr * 10), "07" if m07 else "area")
            sc = gold[key + "_scalars"]
            ref = {"n_pos": gold[key + "_n_pos"], "ap_per_class": gold[key + "_ap_per_class"],
                   "recall_per_class": gold[key + "_recall_per_class"], "map": sc[0], "map_weighted": sc[1], "recall": sc[2],
                   "ap_joint_classes": sc[3]}
            _check(r, ref)
            if not m07 and thr == 0.5:
                np.testing.assert_allclose(r["prec"][4], gold[key + "_prec4"], rtol=0, atol=1e-15, equal_nan=True)
                np.testing.assert_allclose(r["rec"][4], gold[key + "_rec4"], rtol=0, atol=1e-15)
                assert r["prec"][3] is not None and r["rec"][3] is None      # detected but never annotated: no recall curve
                assert len(r["prec"][6]) == 0 and len(r["rec"][6]) == 0      # annotated but never detected


def test_voc_eval_matches_oracle_on_larger_dataset_and_ties():
    from os2d_b200.voc_eval import do_voc_evaluation
    from oracle import voc_oracle as vo

    def arrays(data):
        pb, pl, ps, gb, gl, gd = [], [], [], [], [], []
        for (b, l, s, psize, gt, gtl, gtd, gsize) in data:
            rw, rh = float(gsize[0]) / psize[0], float(gsize[1]) / psize[1]
            scaled = b * rw if rw == rh else b * torch.tensor([rw, rh, rw, rh], dtype=torch.float32)
            pb.append(scaled.numpy()); pl.append(l.numpy()); ps.append(s.numpy())
            gb.append(gt.numpy()); gl.append(gtl.numpy()); gd.append(gtd.numpy())
        return pb, pl, ps, gb, gl, gd

    data = voc_inputs(7, n_images=150, n_labels=40)
    preds, gts = _boxlists(data, "cuda")
    for thr, m07 in ((0.5, False), (0.35, True)):
        _check(do_voc_evaluation(preds, gts, iou_thresh=thr, use_07_metric=m07),
               vo.eval_detection_voc(*arrays(data), iou_thresh=thr, use_07_metric=m07))
    # tied scores: counts and recalls are order-independent, AP only up to the order of the ties
    data = voc_inputs(92, quant=16)
    preds, gts = _boxlists(data, "cuda")
    _check(do_voc_evaluation(preds, gts), vo.eval_detection_voc(*arrays(data)), exact=False)


def test_voc_eval_edge_cases():
    from os2d_b200.voc_eval import do_voc_evaluation
    from os2d_b200.structures import BoxList, FeatureMapSize
    size = FeatureMapSize(w=100, h=100)

    def bl(boxes, labels, scores=None, difficult=None):
        b = BoxList(torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4), size)
        b.add_field("labels", torch.tensor(labels, dtype=torch.int64))
        if scores is not None:
            b.add_field("scores", torch.tensor(scores, dtype=torch.float32))
        if difficult is not None:
            b.add_field("difficult", torch.tensor(difficult, dtype=torch.int64))
        return b

    # no detections at all; ground truth without the "difficult" field
    r = do_voc_evaluation([bl([], [], [])], [bl([[10, 10, 50, 50]], [2])])
    assert r["n_pos"].tolist() == [0, 0, 1] and r["ap_per_class"][2] == 0.0 and np.isnan(r["ap_per_class"][0]) and r["recall"] == 0.0
    # a perfect detection, a duplicate (false positive) and a detection of a difficult box (ignored)
    r = do_voc_evaluation([bl([[10, 10, 50, 50], [11, 11, 50, 50], [60, 60, 90, 90]], [1, 1, 1], [0.9, 0.8, 0.7])],
                          [bl([[10, 10, 50, 50], [60, 60, 90, 90]], [1, 1], difficult=[0, 1])])
    assert r["n_pos"].tolist() == [0, 1] and r["ap_per_class"][1] == 1.0 and r["recall"] == 1.0
    np.testing.assert_allclose(r["prec"][1], [1.0, 0.5, 0.5])
