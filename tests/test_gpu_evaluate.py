"""GPU: the batched-class evaluation iterator (os2d_b200/evaluate.py) against a golden run of the UNMODIFIED reference
iterator (os2d/engine/evaluate.py:177-371, class_batch_size = 1 loop on CPU; tests/golden/make_golden.py)."""
import logging

import numpy as np
import pytest
import torch

from _util import GOLDEN, rel_to_max, TOL
from oracle import head_oracle as ho

pytestmark = pytest.mark.gpu


class _Loader:
    """Duck-typed stand-in for the part of DataloaderOneShotDetection the iterator touches."""

    def __init__(self, class_images, class_ids, pyramids, image_ids, target):
        self.class_images, self.class_ids, self.pyramids, self.image_ids, self.target = class_images, class_ids, pyramids, image_ids, target

    def get_all_class_images(self):
        return self.class_images, [float(im.shape[-1]) / im.shape[-2] for im in self.class_images], self.class_ids

    def make_iterator_for_all_images(self, batch_size, num_random_pyramid_scales=0):
        from os2d_b200.box_coder import make_resize_transform
        transforms = [[make_resize_transform(self.target) for _ in self.pyramids] for _ in self.image_ids]
        yield self.image_ids, self.pyramids, transforms, [self.target] * len(self.image_ids)

    def get_class_ids_for_image_ids(self, image_ids):
        return torch.tensor([self.class_ids[0]])

    def convert_label_ids_global_to_local(self, label_ids, class_ids):
        return torch.tensor([class_ids.index(int(l)) for l in label_ids])


def _inputs():
    g = torch.Generator().manual_seed(55)
    class_images = [torch.randn(1, 3, 64, 80, generator=g), torch.randn(1, 3, 96, 48, generator=g),
                    torch.randn(1, 3, 72, 72, generator=g)]
    pyramids = [torch.randn(2, 3, 96, 128, generator=g), torch.randn(2, 3, 128, 176, generator=g)]
    return class_images, [4, 9, 2], pyramids


def _model(z):
    from os2d_b200.model import Os2dModel
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(int(z["model_seed"]))
    net = Os2dModel(is_cuda=True, backbone_arch="resnet50", merge_branch_parameters=True, use_inverse_geom_model=True,
                    simplify_affine=False)
    chk = float(sum(v.double().abs().sum() for v in net.net_feature_maps.state_dict().values() if v.dtype.is_floating_point))
    assert abs(chk - float(z["backbone_checksum"])) <= 1e-9 * chk, "seeded backbone weights differ from the golden run"
    tn = ho.random_transform_net(6, seed=11, spread=0.005)
    net.os2d_head_creator.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    net.eval()
    return net


def test_eval_iterator_matches_reference_golden():
    from os2d_b200.evaluate import make_iterator_extract_scores_from_images_batched
    from os2d_b200.structures import FeatureMapSize
    z = np.load(GOLDEN + "/eval_iterator.npz")
    net = _model(z)
    class_images, class_ids, pyramids = _inputs()
    loader = _Loader(class_images, class_ids, pyramids, [10, 11], FeatureMapSize(w=352, h=256))
    seen = []
    for rec in make_iterator_extract_scores_from_images_batched(loader, net, logging.getLogger("t"), image_batch_size=2,
                                                                is_cuda=True, class_image_augmentation="horflip"):
        image_id, loc_p, cls_p, pyr, q_sizes, b_class_ids, rev, fm_sizes, corners_p = rec
        seen.append(image_id)
        assert b_class_ids == z["img%d_class_ids" % image_id].tolist()          # 2 views per class: [4,4,9,9,2,2]
        assert [[s.w, s.h] for s in q_sizes] == z["img%d_query_sizes" % image_id].tolist()
        assert len(loc_p) == 2 and len(rev) == 2 and pyr[0].shape == (3, 96, 128)
        for lvl in range(2):
            assert [fm_sizes[lvl].w, fm_sizes[lvl].h] == z["img%d_fm_%d" % (image_id, lvl)].tolist()
            assert loc_p[lvl].shape == z["img%d_loc_%d" % (image_id, lvl)].shape
            # backbone runs in cuDNN fp32 here and in MKL-DNN fp32 in the golden run: well inside the 1e-3 bar
            assert rel_to_max(cls_p[lvl].cpu(), z["img%d_cls_%d" % (image_id, lvl)]) < TOL
            assert rel_to_max(loc_p[lvl].cpu(), z["img%d_loc_%d" % (image_id, lvl)]) < TOL
            assert rel_to_max(corners_p[lvl].cpu(), z["img%d_corners_%d" % (image_id, lvl)]) < TOL
    assert seen == [10, 11]


def test_eval_iterator_label_subset_and_decode():
    """num_random_negative_labels >= 0 searches a label subset (Os2dHead.select); the yielded pyramid feeds decode_pyramid."""
    from os2d_b200.evaluate import make_iterator_extract_scores_from_images_batched
    from os2d_b200.structures import FeatureMapSize
    from os2d_b200.box_coder import Os2dBoxCoder
    z = np.load(GOLDEN + "/eval_iterator.npz")
    net = _model(z)
    class_images, class_ids, pyramids = _inputs()
    loader = _Loader(class_images, class_ids, pyramids, [10, 11], FeatureMapSize(w=352, h=256))
    full = list(make_iterator_extract_scores_from_images_batched(loader, net, logging.getLogger("t"), 2, True))
    torch.manual_seed(0)
    sub = list(make_iterator_extract_scores_from_images_batched(loader, net, logging.getLogger("t"), 2, True,
                                                                num_random_negative_labels=1))
    ids_full, ids_sub = full[0][5], sub[0][5]
    assert len(ids_sub) <= 2 and class_ids[0] in ids_sub
    for k, cid in enumerate(ids_sub):
        j = ids_full.index(cid)
        assert torch.equal(sub[0][2][0][k], full[0][2][0][j])                   # same class plane, bit for bit
    coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, net.os2d_head_creator.box_grid_generator_image_level, net.get_feature_map_size)
    image_id, loc_p, cls_p, pyr, q_sizes, b_class_ids, rev, fm_sizes, corners_p = full[1]
    img_sizes = [FeatureMapSize(img=p) for p in pyr]
    dets = coder.decode_pyramid(loc_p, cls_p, img_sizes, b_class_ids, nms_score_threshold=float(cls_p[0].median()),
                                nms_iou_threshold=0.3, inverse_box_transforms=rev, transform_corners_pyramid=corners_p)
    assert len(dets) > 0 and dets.image_size == FeatureMapSize(w=352, h=256)
    assert set(dets.get_field("labels").tolist()) <= set(class_ids)


class _EvalLoader(_Loader):
    """+ the members the evaluation loop itself uses (os2d/engine/evaluate.py:35-36, 66, 104)."""

    def __init__(self, *a, box_coder=None, annotations=None):
        super().__init__(*a)
        self.box_coder, self.annotations = box_coder, annotations

    def get_name(self):
        return "synthetic-eval"

    def get_eval_scale(self):
        return 1.0

    def get_image_annotation_for_imageid(self, image_id):
        return self.annotations[image_id]


def test_evaluate_loop_metrics_and_detections_dump(tmp_path):
    """os2d_b200.evaluate.evaluate: the inference side of the reference evaluate() (evaluate.py:20-174): losses keys, the
    <dataset>_detections.pth dump format (evaluate.py:136-149) and mAP equal to the CPU oracle on the dumped detections."""
    from os2d_b200.evaluate import evaluate
    from os2d_b200.structures import FeatureMapSize, BoxList
    from os2d_b200.box_coder import Os2dBoxCoder
    from oracle import voc_oracle as vo
    z = np.load(GOLDEN + "/eval_iterator.npz")
    net = _model(z)
    class_images, class_ids, pyramids = _inputs()
    target = FeatureMapSize(w=352, h=256)
    ann = {}
    for k, image_id in enumerate([10, 11]):
        gt = BoxList(torch.tensor([[20.0 + 30 * k, 30.0, 180.0, 170.0], [150.0, 60.0, 330.0, 240.0]]), target)
        gt.add_field("labels", torch.tensor([4, 2 if k == 0 else 9]))
        gt.add_field("difficult", torch.tensor([0, k]))
        ann[image_id] = gt
    coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, net.os2d_head_creator.box_grid_generator_image_level, net.get_feature_map_size)
    loader = _EvalLoader(class_images, class_ids, pyramids, [10, 11], target, box_coder=coder, annotations=ann)
    cfg = {"is_cuda": True,
           "eval": {"batch_size": 2, "class_image_augmentation": "", "nms_iou_threshold": 0.3,
                    "nms_score_threshold": float("-inf"), "mAP_iou_thresholds": [0.5, 0.3]},
           "visualization": {"eval": {"path_to_save_detections": str(tmp_path)}}}
    losses = evaluate(loader, net, cfg)
    assert [k for k in losses] == ["mAP@0.50", "mAPw@0.50", "recall@0.50", "AP_joint_classes@0.50",
                                   "mAP@0.30", "mAPw@0.30", "recall@0.30", "AP_joint_classes@0.30", "eval_time"]
    data = torch.load(str(tmp_path / "synthetic-eval_detections.pth"))
    assert sorted(data.keys()) == sorted(["image_ids", "boxes_xyxy", "labels", "scores", "gt_boxes_xyxy", "gt_labels", "gt_difficults"])
    assert data["image_ids"] == [10, 11] and len(data["boxes_xyxy"]) == 2
    assert all(not t.is_cuda for t in data["boxes_xyxy"] + data["scores"] + data["labels"])
    assert data["boxes_xyxy"][0].shape[0] == data["scores"][0].shape[0] == data["labels"][0].shape[0] > 0
    for thr in (0.5, 0.3):
        ref = vo.eval_detection_voc([b.numpy() for b in data["boxes_xyxy"]], [l.numpy() for l in data["labels"]],
                                    [s.numpy() for s in data["scores"]], [b.numpy() for b in data["gt_boxes_xyxy"]],
                                    [l.numpy() for l in data["gt_labels"]], [d.numpy() for d in data["gt_difficults"]],
                                    iou_thresh=thr)
        assert abs(losses["mAP@{:0.2f}".format(thr)] - ref["map"]) < 1e-9
        assert abs(losses["recall@{:0.2f}".format(thr)] - ref["recall"]) < 1e-9
    with pytest.raises(NotImplementedError):
        evaluate(loader, net, cfg, criterion=object())
