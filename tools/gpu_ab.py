"""In-process A/B timing of two builds of the library (box-to-box and run-to-run variance is a few %, so variants are
compared interleaved in ONE process on the same buffers): A = os2d_b200/libos2d_b200_base.so (tools/build_baseline_lib.sh),
B = os2d_b200/libos2d_b200.so.   python tools/gpu_ab.py [rounds=6] [steps=20] [ENV=VALUE for B only ...]
Environment switches that a library reads once (static getenv) can be given for B only: they are set while B makes its
first calls and unset while A does, e.g. `python tools/gpu_ab.py 6 20 OS2D_B200_CONV3_V2=1` with two copies of one build."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from os2d_b200 import _cabi
from os2d_b200 import head as bh
from os2d_b200.structures import FeatureMapSize
from _synth import seeded_transform_net

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ENV_B = dict(a.split("=", 1) for a in sys.argv[3:])


def open_lib(path):
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _cabi.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


libs = {"A(base)": open_lib(os.path.join(ROOT, "os2d_b200", "libos2d_b200_base.so")),
        "B(new)": open_lib(os.path.join(ROOT, "os2d_b200", "libos2d_b200.so"))}
C, side, D = 100, 80, 1024
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
cms = (torch.randn(C, D, 15, 15, generator=g) * 0.5 + 0.2).relu()
fm = (torch.randn(1, D, side, side, generator=g) * 0.5 + 0.2).relu().to(dev)
tn = seeded_transform_net(6, seed=1, spread=0.005)
heads = {}
acc = {k: {} for k in libs}
tot = {k: [] for k in libs}
outs = {}
with torch.no_grad():
    for name, lib in libs.items():                       # one head (own packed weights) per library
        for k, v in ENV_B.items():
            if name.startswith("B"):
                os.environ[k] = v
            else:
                os.environ.pop(k, None)
        _cabi._lib = lib
        hc = bh.build_os2d_head_creator(False, True, True, FeatureMapSize(w=16, h=16), FeatureMapSize(w=16, h=16))
        hc.aligner.parameter_regressor.load_state_dict(dict(tn), strict=False)
        hc.eval()
        heads[name] = hc.create_os2d_head([cms[i:i + 1].to(dev) for i in range(C)])
        heads[name](fm)                                   # first calls: every static switch of this library is read now
        torch.cuda.synchronize()
    for r in range(R + 1):
        for name, lib in libs.items():
            _cabi._lib = lib
            head = heads[name]
            head.profile_events = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(K):
                o = head(fm)
            e1.record()
            torch.cuda.synchronize()
            outs[name] = o
            if r == 0:
                continue      # warm-up round
            tot[name].append(e0.elapsed_time(e1) / K)
            for st, a, b in head.profile_events:
                acc[name].setdefault(st, []).append(a.elapsed_time(b))
            head.profile_events = None
for name in libs:
    t = sorted(tot[name])
    print("{:8s} step median {:.4f} ms (min {:.4f})  ".format(name, t[len(t) // 2], t[0]) +
          "  ".join("{} {:.4f}".format(st, sorted(v)[len(v) // 2]) for st, v in acc[name].items()), flush=True)
a, b = outs["A(base)"], outs["B(new)"]
for i, nm in ((0, "loc"), (1, "score"), (3, "corners")):
    d = (a[i] - b[i]).abs().max().item()
    print("  {} max|A-B| = {:.3e} (max|A| {:.3e})".format(nm, d, a[i].abs().max().item()))
