"""Measurement of the on-device image pyramid against the reference's CPU path
(PIL resize + ToTensor + Normalize per level, then H2D).   python tools/gpu_pyramid_bench.py > profiles/r01_pyramid.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

from os2d_b200.pyramid import image_pyramid, resize_normalize

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
rng = np.random.default_rng(0)
W, H = 1280, 960
img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
scales = (0.5, 0.625, 0.8, 1, 1.2, 1.4, 1.6)      # configs[3]: 640 ... 2048 px on the long side
pil = Image.fromarray(img, "RGB")


def cpu_path():
    out = []
    for s in scales:
        p = np.asarray(pil.resize((int(W * s), int(H * s)), Image.BILINEAR))
        t = torch.from_numpy(p.copy()).permute(2, 0, 1).to(torch.float32).div(255)
        t = (t - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)
        out.append(t.cuda(non_blocking=False))
    return out


def gpu_path():
    return image_pyramid(img, scales, MEAN, STD)[0]


for fn in (cpu_path, gpu_path):
    fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ref = cpu_path()
torch.cuda.synchronize()
t_cpu = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(10):
    out = gpu_path()
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / 10
same = all(torch.equal(a, b) for a, b in zip(out, ref))
# kernels only: one level 1280x960 -> 2048x1536, device-resident image, events
dimg = torch.from_numpy(img).cuda()
resize_normalize(dimg, 2048, 1536, MEAN, STD)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    resize_normalize(dimg, 2048, 1536, MEAN, STD)
e1.record()
torch.cuda.synchronize()
lvl_ms = e0.elapsed_time(e1) / 20
bytes_lvl = H * W * 3 + 2 * H * 2048 * 3 + 3 * 1536 * 2048 * 4
print(json.dumps({"what": "7-level pyramid of a {}x{} image (scales {}), uint8 image on the host at the start, fp32 CHW levels on the GPU at the end".format(W, H, scales),
                  "cpu_pil_torchvision_ms": t_cpu * 1e3, "device_ms": t_gpu * 1e3, "bit_identical": bool(same),
                  "level_2048x1536_ms_incl_table_upload": lvl_ms, "level_bytes": bytes_lvl,
                  "level_gbs": bytes_lvl / (lvl_ms * 1e-3) / 1e9}))
