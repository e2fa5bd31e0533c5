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
