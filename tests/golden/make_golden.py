"""Generates the golden input/output fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/os2d, imported read-only) on seeded synthetic inputs.  Run in the build container:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors for this path, so these files are the
pin of the oracle: tests/test_oracle_golden.py checks oracle/ against them on any machine, and the GPU parity
tests check the CUDA path against the same files.  The TransformNet weights are regenerated from a seed by
oracle.head_oracle.random_transform_net (checksum stored in the fixture).
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from os2d.modeling.head import build_os2d_head_creator          # noqa: E402  (the reference)
from os2d.modeling.box_coder import Os2dBoxCoder                # noqa: E402
from os2d.structures.feature_map import FeatureMapSize          # noqa: E402
from os2d.structures.transforms import TransformList            # noqa: E402
from os2d.structures.bounding_box import BoxList, nms as ref_nms  # noqa: E402
from oracle import head_oracle as ho                            # noqa: E402

D = 64
TN_SEED = 11
TN_SPREAD = 0.005
VARIANTS = [("affine_inverse", False, True), ("affine", False, False), ("simple", True, False),
            ("simple_inverse", True, True)]


def tn_checksum(tn):
    return float(sum(v.double().abs().sum() for v in tn.values()))


def synth_inputs(seed, B, H, W, sizes):
    g = torch.Generator().manual_seed(seed)
    cms = [(torch.randn(1, D, h, w, generator=g) * 0.5 + 0.2).relu() for (h, w) in sizes]
    fm = (torch.randn(B, D, H, W, generator=g) * 0.5 + 0.2).relu()
    return cms, fm


def ref_head(simple, inverse, tn, cms, fm):
    hc = build_os2d_head_creator(simple, False, inverse, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    sd = dict(tn)
    sd["conv.1.num_batches_tracked"] = torch.tensor(0)
    sd["conv.4.num_batches_tracked"] = torch.tensor(0)
    hc.aligner.parameter_regressor.load_state_dict(sd)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head(cms)
        loc, rec, _, corners = head(fm)
    return hc, head, loc, rec, corners


def make_head_goldens():
    for name, simple, inverse in VARIANTS:
        P = 4 if simple else 6
        tn = ho.random_transform_net(P, seed=TN_SEED, spread=TN_SPREAD)
        sizes = [(15, 15), (9, 17), (20, 8)]
        cms, fm = synth_inputs(100 + P + int(inverse), 2, 13, 11, sizes)
        hc, head, loc, rec, corners = ref_head(simple, inverse, tn, cms, fm)
        out = {"fm": fm.numpy(), "loc": loc.numpy(), "score": rec.numpy(), "corners": corners.numpy(),
               "class_features": head.class_feature_maps.numpy(),
               "tn_seed": np.int64(TN_SEED), "tn_spread": np.float64(TN_SPREAD), "tn_checksum": np.float64(tn_checksum(tn)),
               "simple": np.int64(simple), "inverse": np.int64(inverse)}
        for i, c in enumerate(cms):
            out["class_map_%d" % i] = c.numpy()
        np.savez_compressed(os.path.join(HERE, "head_%s.npz" % name), **out)
        print(name, "loc range", float(loc.min()), float(loc.max()), "score range", float(rec.min()), float(rec.max()))


def make_decode_golden():
    """Two pyramid levels, three class views with a duplicated class id, inverse transforms to a common size."""
    simple, inverse = False, True
    tn = ho.random_transform_net(6, seed=TN_SEED, spread=TN_SPREAD)
    sizes = [(15, 15), (9, 17), (20, 8)]
    img_sizes = [(176, 208), (240, 288)]          # (w, h) of the level images; fm = ceil(./16)
    target = (480, 576)                           # common original image size (both ratios differ per level)
    loc_pyr, cls_pyr, cor_pyr, fm_sizes = [], [], [], []
    hc = None
    for lvl, (iw, ih) in enumerate(img_sizes):
        fw, fh = -(-iw // 16), -(-ih // 16)
        cms, fm = synth_inputs(300 + lvl, 1, fh, fw, sizes)
        if lvl > 0:
            cms, _ = synth_inputs(300, 1, fh, fw, sizes)     # same class maps on every level
        hc, head, loc, rec, corners = ref_head(simple, inverse, tn, cms, fm)
        loc_pyr.append(loc[0].reshape(3, 4, -1))
        cls_pyr.append(rec[0].reshape(3, -1))
        cor_pyr.append(corners[0].reshape(3, 8, -1))
        fm_sizes.append((fw, fh))
    class_ids = [5, 7, 5]
    box_coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level,
                             lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)), do_nms_across_classes=False)
    tgt = FeatureMapSize(w=target[0], h=target[1])
    transforms = []
    for _ in img_sizes:
        t = TransformList()
        t.append(lambda boxes: boxes.resize(tgt))
        transforms.append(t)
    thr = float(torch.cat([c.flatten() for c in cls_pyr]).median())
    res = box_coder.decode_pyramid(loc_pyr, cls_pyr, [FeatureMapSize(w=w, h=h) for (w, h) in img_sizes], class_ids,
                                   nms_score_threshold=thr, nms_iou_threshold=0.3, inverse_box_transforms=transforms,
                                   transform_corners_pyramid=cor_pyr)
    out = {"class_ids": np.array(class_ids), "img_sizes": np.array(img_sizes), "fm_sizes": np.array(fm_sizes),
           "target": np.array(target), "score_thr": np.float64(thr), "iou_thr": np.float64(0.3),
           "boxes": res.bbox_xyxy.numpy(), "scores": res.get_field("scores").numpy(), "labels": res.get_field("labels").numpy(),
           "default_boxes": res.get_field("default_boxes").bbox_xyxy.numpy(),
           "transform_corners": res.get_field("transform_corners").numpy()}
    for lvl in range(len(img_sizes)):
        out["loc_%d" % lvl] = loc_pyr[lvl].numpy()
        out["cls_%d" % lvl] = cls_pyr[lvl].numpy()
        out["corners_%d" % lvl] = cor_pyr[lvl].numpy()
    np.savez_compressed(os.path.join(HERE, "decode_pyramid.npz"), **out)
    print("decode golden:", len(res), "detections, thr", thr)


def make_nms_golden():
    """Chunked NMS path (> 10000 candidates): only the seed and the kept indices are stored."""
    seed, n = 77, 23000
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(n, 2, generator=g) * 600
    wh = torch.rand(n, 2, generator=g) * 120 + 20
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    scores = (torch.rand(n, generator=g) * 64).round() / 64        # many exact ties
    bl = BoxList(boxes, FeatureMapSize(w=800, h=800))
    bl.add_field("scores", scores)
    keep = ref_nms(bl, 0.3)
    np.savez_compressed(os.path.join(HERE, "nms_chunked.npz"), seed=np.int64(seed), n=np.int64(n), keep=keep.numpy(),
                        box_checksum=np.float64(boxes.double().sum()), score_checksum=np.float64(scores.double().sum()))
    print("nms golden: kept", keep.numel(), "of", n)


MODEL_SEED = 7


class FakeLoader:
    """The part of DataloaderOneShotDetection (os2d/data/dataloader.py:146) the evaluation iterator touches."""

    def __init__(self, class_images, class_ids, pyramids, image_ids, target_sizes, ref_types=True):
        self.class_images, self.class_ids = class_images, class_ids
        self.pyramids, self.image_ids, self.target_sizes = pyramids, image_ids, target_sizes

    def get_all_class_images(self):
        ratios = [float(im.shape[-1]) / im.shape[-2] for im in self.class_images]
        return self.class_images, ratios, self.class_ids

    def make_iterator_for_all_images(self, batch_size, num_random_pyramid_scales=0):
        # one batch holding every image; box transforms: one TransformList per (image, level)
        transforms = []
        for tgt in self.target_sizes:
            per_level = []
            for _ in self.pyramids:
                t = TransformList()
                t.append(lambda boxes, tgt=tgt: boxes.resize(tgt))
                per_level.append(t)
            transforms.append(per_level)
        yield self.image_ids, self.pyramids, transforms, self.target_sizes


def eval_iterator_inputs():
    g = torch.Generator().manual_seed(55)
    class_images = [torch.randn(1, 3, 64, 80, generator=g), torch.randn(1, 3, 96, 48, generator=g),
                    torch.randn(1, 3, 72, 72, generator=g)]
    class_ids = [4, 9, 2]
    pyramids = [torch.randn(2, 3, 96, 128, generator=g), torch.randn(2, 3, 128, 176, generator=g)]
    return class_images, class_ids, pyramids


def make_eval_iterator_golden():
    """Runs the UNMODIFIED reference iterator (class_batch_size = 1 loop, evaluate.py:177-371) on CPU with a fake
    dataloader; the model weights are regenerated from MODEL_SEED by the test (checksum stored)."""
    import types
    import logging
    sys.modules.setdefault("os2d.utils.visualization", types.ModuleType("os2d.utils.visualization"))
    from os2d.engine.evaluate import make_iterator_extract_scores_from_images_batched as ref_iterator
    from os2d.modeling.model import Os2dModel
    torch.cuda.synchronize = lambda *a, **k: None          # evaluate.py:312,332 call it unguarded; CPU run here
    torch.manual_seed(MODEL_SEED)
    net = Os2dModel(logger=logging.getLogger("golden"), is_cuda=False, backbone_arch="resnet50",
                    merge_branch_parameters=True, use_inverse_geom_model=True, simplify_affine=False)
    tn = ho.random_transform_net(6, seed=TN_SEED, spread=TN_SPREAD)
    sd = dict(tn)
    sd["conv.1.num_batches_tracked"] = torch.tensor(0)
    sd["conv.4.num_batches_tracked"] = torch.tensor(0)
    net.os2d_head_creator.aligner.parameter_regressor.load_state_dict(sd)
    net.eval()
    backbone_checksum = float(sum(v.double().abs().sum() for k, v in net.net_feature_maps.state_dict().items()
                                  if v.dtype.is_floating_point))
    class_images, class_ids, pyramids = eval_iterator_inputs()
    tgt = FeatureMapSize(w=352, h=256)
    loader = FakeLoader(class_images, class_ids, pyramids, [10, 11], [tgt, tgt])
    out = {"backbone_checksum": np.float64(backbone_checksum), "model_seed": np.int64(MODEL_SEED)}
    torch.set_grad_enabled(False)       # evaluate() runs the iterator under @torch.no_grad() (evaluate.py:20)
    for rec in ref_iterator(loader, net, logging.getLogger("golden"), image_batch_size=2, is_cuda=False,
                            class_image_augmentation="horflip"):
        image_id, loc_p, cls_p, pyr, q_sizes, b_class_ids, rev, fm_sizes, corners_p = rec
        for lvl in range(len(loc_p)):
            out["img%d_loc_%d" % (image_id, lvl)] = loc_p[lvl].numpy()
            out["img%d_cls_%d" % (image_id, lvl)] = cls_p[lvl].numpy()
            out["img%d_corners_%d" % (image_id, lvl)] = corners_p[lvl].numpy()
            out["img%d_fm_%d" % (image_id, lvl)] = np.array([fm_sizes[lvl].w, fm_sizes[lvl].h])
        out["img%d_class_ids" % image_id] = np.array(b_class_ids)
        out["img%d_query_sizes" % image_id] = np.array([[s.w, s.h] for s in q_sizes])
    np.savez_compressed(os.path.join(HERE, "eval_iterator.npz"), **out)
    print("eval iterator golden:", sorted(k for k in out if k.startswith("img10"))[:4], "backbone checksum", backbone_checksum)


# ---------------------------------------------------------------------------------------------------------------------
def singular_params(seed, NB, P, H, W, singular_at):
    """Seeded regressed parameters around identity with exactly singular matrices planted at the given (n, y, x)."""
    g = torch.Generator().manual_seed(seed)
    ident = torch.tensor([1., 0, 0, 0, 1, 0] if P == 6 else [1., 0, 1, 0]).view(1, P, 1, 1)
    p = ident + 0.2 * torch.randn(NB, P, H, W, generator=g)
    kinds6 = [[0., 0, 0.3, 0, 0, -0.2], [1., 2, 0.1, 1, 2, 0.5], [0.5, 0.25, 0.1, 2, 1, 0.5]]
    kinds4 = [[0., 0.3, 1.1, -0.2], [0.9, 0.1, 0., 0.5], [0., 0., 0., 0.]]
    for i, (n, y, x) in enumerate(singular_at):
        p[n, :, y, x] = torch.tensor((kinds6 if P == 6 else kinds4)[i % 3])
    return p


def make_singular_theta_golden():
    """Failure handling of the inverse geometric model (head.py:123-146): the reference's own
    prepare_transform_parameters_for_grid_sampler on parameters with exactly singular matrices.  Case `small`: one chunk
    (< 65535 matrices), everything regularised.  Case `chunks`: 65884 matrices = chunks of 65535 + 349 with singular
    matrices only in the second one; only a window of rows and per-chunk checksums are stored."""
    out = {}
    for name, simple in (("affine", False), ("simple", True)):
        P = 4 if simple else 6
        hc = build_os2d_head_creator(simple, False, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
        # small
        sing = [(0, 1, 2), (1, 3, 4), (1, 0, 0)]
        p = singular_params(21, 2, P, 5, 7, sing)
        with torch.no_grad():
            th = hc.aligner.prepare_transform_parameters_for_grid_sampler(p)
        out["small_{}_theta".format(name)] = th.numpy()
        out["small_{}_singular".format(name)] = np.array(sing, dtype=np.int64)
        # two chunks
        NB, H, W = 2, 182, 181
        sing2 = [(1, 181, 100), (1, 181, 180)]          # flat indices 65803 and 65883: second chunk only
        p2 = singular_params(22, NB, P, H, W, sing2)
        with torch.no_grad():
            th2 = hc.aligner.prepare_transform_parameters_for_grid_sampler(p2).reshape(-1, 6)
        out["chunks_{}_singular".format(name)] = np.array(sing2, dtype=np.int64)
        out["chunks_{}_head".format(name)] = th2[:256].numpy()
        out["chunks_{}_tail".format(name)] = th2[65535 - 128:].numpy()
        out["chunks_{}_sum0".format(name)] = np.float64(th2[:65535].double().sum())
    np.savez_compressed(os.path.join(HERE, "theta_singular.npz"), **out)
    print("singular theta golden:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def voc_inputs(seed, n_images=14, n_labels=9, quant=0):
    """Seeded detections / ground truth per image: (pred boxes, labels, scores, pred image size, gt boxes, labels,
    difficult, gt image size).  Label 3 never has ground truth, label 6 is never detected, image 2 has no detections,
    image 5 no ground truth; every third image predicts at another resolution (exercises BoxList.resize).
    quant > 0 quantises the scores (ties)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_images):
        W, H = 640 + 32 * (i % 3), 480 + 16 * (i % 4)
        ng = 0 if i == 5 else int(torch.randint(1, 9, (1,), generator=g))
        gxy = torch.rand(ng, 2, generator=g) * torch.tensor([W - 160.0, H - 120.0])
        gwh = 30 + torch.rand(ng, 2, generator=g) * 120
        gt = torch.cat([gxy, gxy + gwh], 1)
        gl = torch.randint(0, n_labels, (ng,), generator=g)
        gl[gl == 3] = 4
        gd = (torch.rand(ng, generator=g) < 0.25).to(torch.int64)
        nd = 0 if i == 2 else int(torch.randint(5, 60, (1,), generator=g))
        # half of the detections are jittered copies of ground-truth boxes (several per box: duplicates), half random
        src = torch.randint(0, max(ng, 1), (nd,), generator=g)
        jit = (torch.rand(nd, 4, generator=g) - 0.5) * 40
        near = (gt[src] + jit) if ng > 0 else torch.zeros(nd, 4)
        rxy = torch.rand(nd, 2, generator=g) * torch.tensor([W - 100.0, H - 100.0])
        rnd = torch.cat([rxy, rxy + 20 + torch.rand(nd, 2, generator=g) * 150], 1)
        use_near = (torch.rand(nd, generator=g) < 0.6) & (ng > 0)
        pb = torch.where(use_near[:, None], near, rnd)
        pb = torch.stack([torch.minimum(pb[:, 0], pb[:, 2]), torch.minimum(pb[:, 1], pb[:, 3]),
                          torch.maximum(pb[:, 0], pb[:, 2]), torch.maximum(pb[:, 1], pb[:, 3])], 1)
        pl = torch.where(use_near & (torch.rand(nd, generator=g) < 0.8), gl[src] if ng > 0 else torch.zeros(nd, dtype=torch.int64),
                         torch.randint(0, n_labels, (nd,), generator=g))
        pl[pl == 6] = 7
        ps = torch.rand(nd, generator=g)
        if quant:
            ps = (ps * quant).floor() / quant
        sx = 0.5 if i % 3 == 1 else 1.0             # predictions made on a smaller image: equal ratios (one multiply in
        sy = (0.5 if i % 2 else 0.25) if i % 3 == 1 else 1.0   # BoxList.resize) or different ratios (per-axis branch)
        psize = (int(W * sx), int(H * sy))
        out.append((pb * torch.tensor([sx, sy, sx, sy]), pl, ps, psize, gt, gl, gd, (W, H)))
    return out


def make_voc_golden():
    """do_voc_evaluation of the reference (voc_eval.py) on the seeded detections above; distinct scores and tied scores,
    IoU thresholds 0.5 and 0.3 (fp32 comparison), both AP definitions."""
    from os2d.data.voc_eval import do_voc_evaluation
    out = {}
    for tag, seed, quant in (("distinct", 91, 0), ("ties", 92, 16)):
        data = voc_inputs(seed, quant=quant)
        preds, gts = [], []
        for (pb, pl, ps, psize, gt, gl, gd, gsize) in data:
            b = BoxList(pb, FeatureMapSize(w=psize[0], h=psize[1]), mode="xyxy")
            b.add_field("labels", pl)
            b.add_field("scores", ps)
            preds.append(b)
            t = BoxList(gt, FeatureMapSize(w=gsize[0], h=gsize[1]), mode="xyxy")
            t.add_field("labels", gl)
            t.add_field("difficult", gd)
            gts.append(t)
        for thr in (0.5, 0.3):
            for m07 in (False, True):
                r = do_voc_evaluation(preds, gts, iou_thresh=thr, use_07_metric=m07)
                key = "{}_thr{}_{}".format(tag, int(thr * 10), "07" if m07 else "area")
                for k in ("ap_per_class", "recall_per_class", "n_pos"):
                    out[key + "_" + k] = np.asarray(r[k], dtype=np.float64)
                out[key + "_scalars"] = np.array([r["map"], r["map_weighted"], r["recall"], r["ap_joint_classes"]], dtype=np.float64)
                if not m07 and thr == 0.5:
                    out[key + "_prec4"] = np.asarray(r["prec"][4], dtype=np.float64)
                    out[key + "_rec4"] = np.asarray(r["rec"][4], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "voc_eval.npz"), **out)
    print("voc golden:", {k: out[k].tolist() for k in out if k.endswith("thr5_area_scalars")})


if __name__ == "__main__":
    make_head_goldens()
    make_decode_golden()
    make_nms_golden()
    make_eval_iterator_golden()
    make_singular_theta_golden()
    make_voc_golden()
