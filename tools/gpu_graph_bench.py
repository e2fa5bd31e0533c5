"""Launch-bound small workload (configs[0]: 512 px, 4 classes, V2): eager head calls vs CUDA-graph replay.
    python tools/gpu_graph_bench.py > profiles/r01_graph_cfg1.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from os2d_b200 import head as bh, GraphedHead
from os2d_b200.structures import FeatureMapSize
from _synth import seeded_transform_net

res = {}
for name, C, side in (("cfg1_512px_c4", 4, 32), ("640px_c16", 16, 40), ("cfg2_1280px_c100", 100, 80)):
    g = torch.Generator().manual_seed(0)
    cms = (torch.randn(C, 1024, 15, 15, generator=g) * 0.5 + 0.2).relu().cuda()
    fm = (torch.randn(1, 1024, side, side, generator=g) * 0.5 + 0.2).relu().cuda()
    tn = seeded_transform_net(6, seed=1, spread=0.005)
    hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
    hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
    hc.eval()
    with torch.no_grad():
        head = hc.create_os2d_head([cms[i:i + 1] for i in range(C)])
        graphed = GraphedHead(head, fm)

        def timed(fn, n=200):
            for _ in range(10):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        eager = timed(lambda: head(fm))
        replay = timed(lambda: graphed(fm))
    res[name] = {"eager_ms": eager, "graph_replay_ms": replay, "classes_per_s_eager": C / eager * 1e3,
                 "classes_per_s_graph": C / replay * 1e3}
print(json.dumps({"what": "Os2dHead.forward eager vs GraphedHead replay, device-resident input, 200 calls", "results": res}))
