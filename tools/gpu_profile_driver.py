"""One pass of every kernel on the path at configs[1] size inside a cudaProfilerStart/Stop range (ncu --profile-from-start off):
class pack, image pack, correlation, conv1..3, resample, fused decode+NMS, gather.  usage: python tools/gpu_profile_driver.py [C] [side]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from os2d_b200 import head as bh  # noqa: E402
from os2d_b200.box_coder import Os2dBoxCoder  # noqa: E402
from os2d_b200.structures import FeatureMapSize  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100
side = int(sys.argv[2]) if len(sys.argv) > 2 else 80
N = side * side
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
fm = (torch.randn(1, 1024, side, side, generator=g) * 0.5 + 0.2).relu().to(dev)
cls = bench.synth_classes(C, 1234).to(dev)
hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
hc.aligner.parameter_regressor.load_state_dict(dict(bench.seeded_transform_net(6, seed=1, spread=0.005)), strict=False)
hc.eval()
coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level, lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
img = FeatureMapSize(w=side * 16, h=side * 16)


def one_pass():
    with torch.no_grad():
        head = hc.create_os2d_head([cls[i:i + 1] for i in range(C)])
        loc, score, _, corners = head(fm)
        return coder.decode_pyramid([loc[0].view(C, 4, N)], [score[0].view(C, N)], [img], list(range(C)),
                                    nms_score_threshold=float("-inf"), nms_iou_threshold=0.3,
                                    transform_corners_pyramid=[corners[0].view(C, 8, N)])


for _ in range(2):
    dets = one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.start()
dets = one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("detections", len(dets))
