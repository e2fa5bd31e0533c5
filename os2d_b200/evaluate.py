"""Batched-class evaluation iterator ("next" row 1).

Replacement for ``make_iterator_extract_scores_from_images_batched`` of the reference
(os2d/engine/evaluate.py:177-371) with the same arguments and the same yield tuple.  The reference hard-codes
``class_batch_size = 1`` (evaluate.py:226): C single-class heads and C x levels head calls per image batch, so an
unmodified ``main.py`` sees tiny GEMMs.  Here ALL class views form ONE multi-class ``Os2dHead`` (the packed fp16 class
operand is built once) and every pyramid level is one head call; results stay on the device.  One multi-class head equals
C single-class heads bit for bit (every (image, class) plane is computed independently, tests/test_gpu_head.py), so the
yielded tensors are identical to what the per-class loop over this package's heads would produce.
"""
import torch

from .structures import FeatureMapSize


def _class_views(im, class_image_augmentation):
    """Augmented views of one class image [3,h,w], in the reference's order (evaluate.py:241-269)."""
    if not class_image_augmentation:
        return [im]
    if class_image_augmentation == "rotation90":
        im90 = im.rot90(1, [1, 2])
        im180 = im90.rot90(1, [1, 2])
        im270 = im180.rot90(1, [1, 2])
        return [im, im90, im180, im270]
    if class_image_augmentation == "horflip":
        return [im, im.flip(2)]
    if class_image_augmentation == "horflip_rotation90":
        im90 = im.rot90(1, [1, 2])
        im180 = im90.rot90(1, [1, 2])
        im270 = im180.rot90(1, [1, 2])
        return [im, im90, im180, im270, im.flip(2), im90.flip(2), im180.flip(2), im270.flip(2)]
    raise RuntimeError("Unknown value of class_image_augmentation: {}".format(class_image_augmentation))


@torch.no_grad()
def make_iterator_extract_scores_from_images_batched(dataloader, net, logger, image_batch_size, is_cuda,
                                                     num_random_pyramid_scales=0, num_random_negative_labels=-1,
                                                     class_image_augmentation=""):
    """Same contract as the reference generator (evaluate.py:177-371).  Yields, per image:
    (image_id, loc scores per level [labels,4,anchors], class scores per level [labels,anchors], image pyramid,
     query image sizes, class ids, box reverse transforms, feature-map sizes per level, transform corners per level
     [labels,8,anchors])."""
    if not is_cuda:
        raise RuntimeError("os2d_b200 requires CUDA (no CPU path)")
    logger.info("Extracting scores from all images")
    class_images, class_aspect_ratios, class_ids = dataloader.get_all_class_images()
    num_classes = len(class_images)
    assert len(class_aspect_ratios) == num_classes and len(class_ids) == num_classes
    query_img_sizes = [FeatureMapSize(img=img) for img in class_images]

    # ---- all class views -> ONE head (class-side operand packed once) ----
    views = []
    num_class_views = 1
    for im in class_images:
        v = _class_views(im.squeeze(0).cuda(), class_image_augmentation)
        num_class_views = len(v)
        views.extend(v)
    logger.info("Extracting weights from {0} classes{1}".format(
        num_classes, " with {} augmentation".format(class_image_augmentation) if class_image_augmentation else ""))
    class_feature_maps = net.net_label_features(views)
    head_all = net.os2d_head_creator.create_os2d_head(class_feature_maps)
    num_views = len(views)

    iterator_batches = dataloader.make_iterator_for_all_images(image_batch_size,
                                                               num_random_pyramid_scales=num_random_pyramid_scales)
    for batch_ids, pyramids_batch, box_transforms_batch, initial_img_size_batch in iterator_batches:
        # labels searched in this batch (evaluate.py:282-295)
        if num_random_negative_labels >= 0:
            neg_labels = torch.randperm(num_views)[:num_random_negative_labels]
            pos_labels = dataloader.get_class_ids_for_image_ids(batch_ids)
            pos_labels = dataloader.convert_label_ids_global_to_local(pos_labels, class_ids)
            batch_labels_local = torch.cat([neg_labels, pos_labels], 0).unique()
            head = head_all.select(batch_labels_local)
        else:
            batch_labels_local = torch.arange(num_views)
            head = head_all
        batch_class_ids = [class_ids[int(l) // num_class_views] for l in batch_labels_local]
        batch_query_img_sizes = [query_img_sizes[int(l) // num_class_views] for l in batch_labels_local]

        batch_images_pyramid, loc_scores, class_scores, fm_sizes, transform_corners = [], [], [], [], []
        for batch_images in pyramids_batch:
            batch_images = batch_images.cuda()
            if getattr(net, "use_packed_feature_maps", False):
                feature_maps = net.net_feature_maps.forward_packed(batch_images)   # fp16 operand straight from the backbone
            else:
                feature_maps = net.net_feature_maps(batch_images)
            loc_s, class_s, _, fm_size, corners = net(class_head=head, feature_maps=feature_maps)
            loc_scores.append(loc_s)              # [B, labels, 4, anchors]
            class_scores.append(class_s)          # [B, labels, anchors]
            transform_corners.append(corners)     # [B, labels, 8, anchors]
            fm_sizes.append(fm_size)
            del feature_maps
            batch_images_pyramid.append(batch_images)

        for i_image_in_batch, image_id in enumerate(batch_ids):
            yield (image_id,
                   [s[i_image_in_batch] for s in loc_scores],
                   [s[i_image_in_batch] for s in class_scores],
                   [p[i_image_in_batch] for p in batch_images_pyramid],
                   batch_query_img_sizes, batch_class_ids, box_transforms_batch[i_image_in_batch],
                   list(fm_sizes),
                   [s[i_image_in_batch] for s in transform_corners])


def _cfg_get(cfg, path, default=None):
    """cfg.a.b.c for yacs-like nodes, plain namespaces and nested dicts; ``default`` when a component is missing."""
    node = cfg
    for name in path.split("."):
        if node is None:
            return default
        node = node.get(name, None) if isinstance(node, dict) else getattr(node, name, None)
    return default if node is None else node


@torch.no_grad()
def evaluate(dataloader, net, cfg, criterion=None, print_per_class_results=False):
    """Evaluation of a model on one dataset: drop-in for the inference side of the reference's ``evaluate``
    (os2d/engine/evaluate.py:20-174) - same arguments, the same ``losses`` keys (mAP@t, mAPw@t, recall@t,
    AP_joint_classes@t per cfg.eval.mAP_iou_thresholds, eval_time) and the same ``<dataset>_detections.pth`` dump
    (evaluate.py:136-149: image_ids, boxes_xyxy / labels / scores per image, gt_boxes_xyxy / gt_labels / gt_difficults).

    Scores come from the batched-class iterator above, decoding + NMS from ``dataloader.box_coder.decode_pyramid`` (this
    package's coder or the reference's hooked one) and the mAP from the on-device ``do_voc_evaluation``; detections stay on
    the GPU until the dump.  ``criterion`` (the training objective evaluated on the targets, evaluate.py:62-101) needs the
    training-side target encoding, which is outside the hot path: passing one raises."""
    import logging
    import os
    import time
    from collections import OrderedDict

    from .voc_eval import do_voc_evaluation
    if criterion is not None:
        raise NotImplementedError("os2d_b200.evaluate: the loss metrics of the training objective (criterion) are outside the "
                                  "inference hot path; call with criterion=None")
    logger = logging.getLogger("OS2D.evaluate")
    dataset_name = dataloader.get_name()
    logger.info("Starting to eval on {0}, scale {1}".format(dataset_name, dataloader.get_eval_scale()))
    t_start_eval = time.time()
    net.eval()
    iterator = make_iterator_extract_scores_from_images_batched(
        dataloader, net, logger, image_batch_size=_cfg_get(cfg, "eval.batch_size", 1), is_cuda=_cfg_get(cfg, "is_cuda", True),
        class_image_augmentation=_cfg_get(cfg, "eval.class_image_augmentation", ""))
    boxes, gt_boxes, image_ids = [], [], []
    losses = OrderedDict()
    for (image_id, loc_pyramid, cls_pyramid, image_pyramid, _query_sizes, class_ids, box_reverse_transform, _fm_sizes,
         corners_pyramid) in iterator:
        image_ids.append(image_id)
        gt_boxes.append(dataloader.get_image_annotation_for_imageid(image_id))
        img_size_pyramid = [FeatureMapSize(img=img) for img in image_pyramid]
        dets = dataloader.box_coder.decode_pyramid(loc_pyramid, cls_pyramid, img_size_pyramid, class_ids,
                                                   nms_iou_threshold=_cfg_get(cfg, "eval.nms_iou_threshold", 0.3),
                                                   nms_score_threshold=_cfg_get(cfg, "eval.nms_score_threshold", float("-inf")),
                                                   inverse_box_transforms=box_reverse_transform,
                                                   transform_corners_pyramid=corners_pyramid)
        boxes.append(dets)

    path_to_save_detections = _cfg_get(cfg, "visualization.eval.path_to_save_detections", "")
    if path_to_save_detections:
        data = {"image_ids": image_ids,
                "boxes_xyxy": [bb.bbox_xyxy.cpu() for bb in boxes],
                "labels": [bb.get_field("labels").cpu() for bb in boxes],
                "scores": [bb.get_field("scores").cpu() for bb in boxes],
                "gt_boxes_xyxy": [bb.bbox_xyxy.cpu() for bb in gt_boxes],
                "gt_labels": [bb.get_field("labels").cpu() for bb in gt_boxes],
                "gt_difficults": [bb.get_field("difficult").cpu() for bb in gt_boxes]}
        os.makedirs(path_to_save_detections, exist_ok=True)
        torch.save(data, os.path.join(path_to_save_detections, dataset_name + "_detections.pth"))

    for thr in _cfg_get(cfg, "eval.mAP_iou_thresholds", [0.5]):
        logger.info("Evaluating at IoU th {:0.2f}".format(thr))
        ap_data = do_voc_evaluation(boxes, gt_boxes, iou_thresh=thr, use_07_metric=False)
        losses["mAP@{:0.2f}".format(thr)] = ap_data["map"]
        losses["mAPw@{:0.2f}".format(thr)] = ap_data["map_weighted"]
        losses["recall@{:0.2f}".format(thr)] = ap_data["recall"]
        losses["AP_joint_classes@{:0.2f}".format(thr)] = ap_data["ap_joint_classes"]
        if print_per_class_results:
            for i_class, (ap, recall, n_pos) in enumerate(zip(ap_data["ap_per_class"], ap_data["recall_per_class"], ap_data["n_pos"])):
                if ap == ap:        # not NaN
                    logger.info("Class {0}, AP {1:0.4f}, #obj {2}, recall {3:0.4f}".format(i_class, ap, n_pos, recall))
    losses["eval_time"] = time.time() - t_start_eval
    logger.info("Evaluated on {0}, scale {1}".format(dataset_name, dataloader.get_eval_scale()))
    return losses
