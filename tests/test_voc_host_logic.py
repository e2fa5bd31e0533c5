"""CPU: the host / tensor part of the on-device detection evaluation (greedy flags, segmented curves, both AP definitions,
result dict of os2d_b200.voc_eval) with the matching step emulated in numpy, against the numpy oracle (which is pinned to
the reference by tests/golden/voc_eval.npz).  The matching kernel itself is covered on the GPU (tests/test_gpu_voc.py)."""
import numpy as np
import pytest
import torch

from _util import voc_inputs
from oracle import voc_oracle as vo


def _arrays(data):
    pb, pl, ps, gb, gl, gd = [], [], [], [], [], []
    for (b, l, s, psize, gt, gtl, gtd, gsize) in data:
        rw, rh = float(gsize[0]) / psize[0], float(gsize[1]) / psize[1]
        scaled = b * rw if rw == rh else b * torch.tensor([rw, rh, rw, rh], dtype=torch.float32)
        pb.append(scaled.numpy()); pl.append(l.numpy()); ps.append(s.numpy())
        gb.append(gt.numpy()); gl.append(gtl.numpy()); gd.append(gtd.numpy())
    return pb, pl, ps, gb, gl, gd


def _emulated_matching(pb, pl, gb, gl, thr):
    """What csrc/voc.cu computes: global ground-truth index of the best box of the same image and label, or -1."""
    out, off = [], 0
    for i in range(len(pb)):
        gi = np.full(len(pb[i]), -1, dtype=np.int64)
        for l in np.unique(pl[i]):
            pm = pl[i] == l
            gm = np.where(gl[i] == l)[0]
            if gm.size == 0 or pm.sum() == 0:
                continue
            iou = vo.box_iou_plus_one(pb[i][pm], gb[i][gm])
            g = gm[iou.argmax(1)] + off
            g[iou.max(1) < np.float32(thr)] = -1
            gi[pm] = g
        out.append(gi)
        off += len(gb[i])
    return np.concatenate(out)


@pytest.mark.parametrize("seed,n_images,n_labels,thr,m07", [(91, 14, 9, 0.5, False), (91, 14, 9, 0.3, True),
                                                            (7, 150, 40, 0.5, False), (7, 150, 40, 0.35, True),
                                                            (3, 400, 200, 0.5, False)])
def test_curves_and_ap_match_oracle(seed, n_images, n_labels, thr, m07):
    from os2d_b200 import voc_eval as ve
    pb, pl, ps, gb, gl, gd = _arrays(voc_inputs(seed, n_images=n_images, n_labels=n_labels))
    ref = vo.eval_detection_voc(pb, pl, ps, gb, gl, gd, iou_thresh=thr, use_07_metric=m07)
    det_l, det_s = torch.from_numpy(np.concatenate(pl)), torch.from_numpy(np.concatenate(ps))
    gt_l, gt_d = torch.from_numpy(np.concatenate(gl)), torch.from_numpy(np.concatenate(gd)).bool()
    gi = torch.from_numpy(_emulated_matching(pb, pl, gb, gl, thr))
    L = int(max(det_l.max(), gt_l.max())) + 1
    r = ve._evaluate_from_matches(det_l, det_s, gi, gt_l, gt_d, L, m07)
    for k in ("ap_per_class", "recall_per_class", "n_pos"):
        np.testing.assert_allclose(r[k], ref[k], rtol=0, atol=1e-12, equal_nan=True)
    for k in ("map", "map_weighted", "recall", "ap_joint_classes"):
        assert abs(r[k] - ref[k]) < 1e-12, (k, r[k], ref[k])
    for l in range(L):
        assert (r["prec"][l] is None) == (ref["prec"][l] is None) and (r["rec"][l] is None) == (ref["rec"][l] is None)
        if ref["rec"][l] is not None:
            np.testing.assert_allclose(r["rec"][l], ref["rec"][l], rtol=0, atol=1e-15)
            np.testing.assert_allclose(r["prec"][l], ref["prec"][l], rtol=0, atol=1e-15, equal_nan=True)


def test_requires_cuda_for_the_public_entry_point():
    from os2d_b200.voc_eval import do_voc_evaluation
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        do_voc_evaluation([], [])
