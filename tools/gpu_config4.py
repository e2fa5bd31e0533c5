"""BASELINE.json configs[3] end to end on N GPUs (torchrun): 7-scale pyramid (0.5-1.6) of a 1280 px image = feature maps
40..128 (52 740 anchors per class), 306 classes, V1 head (simplified affine P = 4, no inverse), labels sharded over the ranks,
decode + chunked NMS with the reference defaults (score threshold -inf, IoU 0.3) on each rank's labels, survivors gathered.
Reference: os2d/engine/evaluate.py:306-327 (per-level head calls), os2d/config.py:194 (scales), box_coder.py:448-536.
Prints one JSON line on rank 0.   python -m torch.distributed.run --nproc-per-node 4 ... tools/gpu_config4.py [classes] [images]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from os2d_b200 import dist as bd  # noqa: E402
from os2d_b200 import head as bh  # noqa: E402
from os2d_b200.box_coder import Os2dBoxCoder, make_resize_transform  # noqa: E402
from os2d_b200.structures import FeatureMapSize  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 306
IMAGES = int(sys.argv[2]) if len(sys.argv) > 2 else 4
SIDES = [40, 50, 64, 80, 96, 112, 128]          # ceil(1280 * s / 16) for s in (0.5, 0.625, 0.8, 1, 1.2, 1.4, 1.6)

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

hc = bh.build_os2d_head_creator(True, True, False, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))     # V1 head
hc.aligner.parameter_regressor.load_state_dict(dict(bench.seeded_transform_net(4, seed=1, spread=0.005)), strict=False)
hc.eval()
coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level, lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
lo, hi = bd.shard_bounds(C, world, rank)
own = bench.synth_classes(hi - lo, 777 + rank).to(dev)
maps = [None] * C
for i in range(hi - lo):
    maps[lo + i] = own[i:i + 1]
g = torch.Generator().manual_seed(5)
pyramid = [(torch.randn(1, 1024, s, s, generator=g) * 0.5 + 0.2).relu().to(dev) for s in SIDES]
sizes = [FeatureMapSize(w=16 * s, h=16 * s) for s in SIDES]
inverse = [make_resize_transform(FeatureMapSize(w=1280, h=1280)) for _ in SIDES]
kw = dict(nms_score_threshold=float("-inf"), nms_iou_threshold=0.3, inverse_box_transforms=inverse)

with torch.no_grad():
    det = bd.ClassShardedDetector(maps, list(range(C)), hc.create_os2d_head, coder)
    dets = det(pyramid, sizes, **kw)                      # warm-up (allocations, symmetric-memory / NCCL set-up)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(IMAGES):
        dets = det(pyramid, sizes, **kw)
    e1.record()
    torch.cuda.synchronize()
    # the head alone over the 7 levels (this rank's classes)
    h0.record()
    for _ in range(IMAGES):
        for fm in pyramid:
            det.head(fm)
    h1.record()
    torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / IMAGES, h0.elapsed_time(h1) / IMAGES], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    anchors = sum(s * s for s in SIDES)
    print(json.dumps({"workload": "configs[3]: 7-scale pyramid of a 1280 px image (feature maps {}), {} classes, V1 simplified-affine "
                                  "head, label-sharded over {} GPU(s), decode + chunked NMS (threshold -inf) + gather of the survivors"
                                  .format(SIDES, C, world),
                      "n_gpus": world, "classes": C, "anchors_per_class": anchors, "images_timed": IMAGES,
                      "ms_per_image": float(ms[0]), "classes_per_s": C / (float(ms[0]) * 1e-3),
                      "class_level_pairs_per_s": 7 * C / (float(ms[0]) * 1e-3),
                      "head_only_ms_per_image": float(ms[1]), "head_only_classes_per_s": C / (float(ms[1]) * 1e-3),
                      "detections": len(dets), "candidates_per_class": anchors}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
