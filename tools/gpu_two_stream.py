"""Experiment: throughput of the head when consecutive images run on alternating CUDA streams (the small non-persistent
kernels and the tails of the persistent ones overlap with the next image's kernels) vs one stream.
    python tools/gpu_two_stream.py [streams=2] [steps=40]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from os2d_b200 import head as bh
from os2d_b200.structures import FeatureMapSize
from _synth import seeded_transform_net

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
C, side, D = 100, 80, 1024
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
cms = (torch.randn(C, D, 15, 15, generator=g) * 0.5 + 0.2).relu()
fm = (torch.randn(1, D, side, side, generator=g) * 0.5 + 0.2).relu().to(dev)
tn = seeded_transform_net(6, seed=1, spread=0.005)
hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
hc.eval()
with torch.no_grad():
    head = hc.create_os2d_head([cms[i:i + 1].to(dev) for i in range(C)])
    ref = head(fm)
    torch.cuda.synchronize()

    def run(nstreams, steps):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        main = torch.cuda.current_stream()
        outs = [None] * nstreams
        for s in streams:
            s.wait_stream(main)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            s = streams[i % nstreams]
            with torch.cuda.stream(s):
                outs[i % nstreams] = head(fm)
        for s in streams:
            main.wait_stream(s)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, outs

    for n in (1, S, 1, S, 3):
        run(n, 6)
        ms, outs = run(n, K)
        same = all(torch.equal(o[1], ref[1]) and torch.equal(o[0], ref[0]) for o in outs if o is not None)
        print("streams {}: {:.4f} ms/step  ({:.0f} classes/s)  outputs identical to 1-stream: {}".format(n, ms, C / ms * 1e3, same), flush=True)
