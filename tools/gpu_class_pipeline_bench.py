"""Measurement of the class-side pipeline: C class images of mixed sizes -> backbone class branch
-> resize to 15x15 + L2 norm + fp16 GEMM operand.  Compares the reference's structure (one backbone call and one resize per
class image, model.py:80-88 / head.py:241-259) with the size-batched branch + one ragged pack launch.
    python tools/gpu_class_pipeline_bench.py [C=100] > profiles/r01_class_pipeline.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from os2d_b200 import head as bh
from os2d_b200.model import Os2dModel

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100
torch.manual_seed(0)
net = Os2dModel(is_cuda=True, backbone_arch="resnet50", use_inverse_geom_model=True, simplify_affine=False)
net.eval()
g = torch.Generator().manual_seed(1)
shapes = [(240, 240), (192, 304), (320, 176), (240, 240)]
images = [torch.randn(3, *shapes[i % 4], generator=g).cuda() for i in range(C)]


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


with torch.no_grad():
    extractor = net.net_label_features.net_class_features
    loop_maps = [extractor(im.unsqueeze(0)) for im in images]
    batched_maps = net.net_label_features(images)

    t_backbone_loop = timed(lambda: [extractor(im.unsqueeze(0)) for im in images])
    t_backbone_batched = timed(lambda: net.net_label_features(images))
    t_pack_per_class = timed(lambda: [bh._prepare_class_operands([m]) for m in loop_maps])
    t_pack_ragged = timed(lambda: bh._prepare_class_operands(batched_maps))
    t_total_loop = timed(lambda: [bh._prepare_class_operands([extractor(im.unsqueeze(0))]) for im in images])
    t_total_new = timed(lambda: net.os2d_head_creator.create_os2d_head(net.net_label_features(images)))
    # the kernel alone: descriptors built once, 20 back-to-back launches through the C ABI between two events
    from os2d_b200 import _cabi
    lib = _cabi.load()
    maps = [m.contiguous() for m in batched_maps]
    ptrs = torch.tensor([m.data_ptr() for m in maps], dtype=torch.int64).cuda()
    hw = torch.tensor([[m.size(2), m.size(3)] for m in maps], dtype=torch.int32).cuda()
    cf32 = torch.empty(C, 1024, 15, 15, device="cuda")
    packed = torch.empty(C, 240, 1024, dtype=torch.float16, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(20):
            _cabi.check(lib.os2d_pack_class_features_ragged(_cabi.ptr(ptrs), _cabi.ptr(hw), C, 1024, 1, _cabi.ptr(cf32),
                                                            _cabi.ptr(packed), _cabi.stream_ptr()), "pack")
        ev1.record()
        torch.cuda.synchronize()
    in_bytes = sum(m.numel() for m in batched_maps) * 4
    out_bytes = C * (1024 * 225 * 4 + 240 * 1024 * 2)
print(json.dumps({
    "what": "class-side pipeline, {} class images of sizes {} (ResNet-50 C4 class branch, random init, eval)".format(C, shapes[:3]),
    "backbone_ms": {"per_image_loop": t_backbone_loop, "size_batched": t_backbone_batched},
    "resize_norm_pack_ms": {"per_class_launches": t_pack_per_class, "one_ragged_launch": t_pack_ragged,
                            "ragged_kernel_device_ms": ev0.elapsed_time(ev1) / 20,
                            "ragged_kernel_gbs": (in_bytes + out_bytes) / (ev0.elapsed_time(ev1) / 20 * 1e-3) / 1e9,
                            "algorithmic_bytes": in_bytes + out_bytes},
    "classes_per_s": {"reference_structure": C / t_total_loop * 1e3, "this_pipeline": C / t_total_new * 1e3},
    "note": "wall clock incl. host launch overhead, 5 repetitions after warm-up; backbone = stock torch/cuDNN (out of the hot path)",
}))
