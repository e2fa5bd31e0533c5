"""Times decode + NMS (second figure of the metric) on the head outputs of the bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from os2d_b200 import head as bh
from os2d_b200.structures import FeatureMapSize
from os2d_b200.box_coder import Os2dBoxCoder
from _synth import seeded_transform_net

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = 80
g = torch.Generator().manual_seed(0)
cms = (torch.randn(C, 1024, 15, 15, generator=g) * 0.5 + 0.2).relu().cuda()
fm = (torch.randn(1, 1024, S, S, generator=g) * 0.5 + 0.2).relu().cuda()
tn = seeded_transform_net(6, seed=1, spread=0.005)
hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
hc.eval()
coder = Os2dBoxCoder(0.5, 0.1, 0.8, 0.4, hc.box_grid_generator_image_level, lambda s: FeatureMapSize(w=-(-s.w // 16), h=-(-s.h // 16)))
with torch.no_grad():
    head = hc.create_os2d_head([cms[i:i + 1] for i in range(C)])
    loc, score, _, corners = head(fm)
torch.cuda.synchronize()
img = FeatureMapSize(w=S * 16, h=S * 16)
for thr in (float("-inf"), float(score.median()), float(score.flatten().kthvalue(int(score.numel() * 0.99)).values)):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dets = coder.decode_pyramid([loc[0].reshape(C, 4, -1)], [score[0].reshape(C, -1)], [img], list(range(C)),
                                    nms_score_threshold=thr, nms_iou_threshold=0.3, transform_corners_pyramid=[corners[0].reshape(C, 8, -1)])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("thr {:.4f}: decode+NMS {:.2f} ms for {} classes -> {} detections ({:.0f} classes/s)".format(thr, dt * 1e3, C, len(dets), C / dt), flush=True)

if os.environ.get("PP_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    thr = float("-inf")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        dets = coder.decode_pyramid([loc[0].reshape(C, 4, -1)], [score[0].reshape(C, -1)], [img], list(range(C)),
                                    nms_score_threshold=thr, nms_iou_threshold=0.3, transform_corners_pyramid=[corners[0].reshape(C, 8, -1)])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
