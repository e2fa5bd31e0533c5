"""Measurement of the on-device detection evaluation against the numpy restatement of the
reference's voc_eval.py on the host.   python tests/tools/gpu_voc_bench.py > profiles/r01_voc_eval.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from _util import voc_inputs
from oracle import voc_oracle as vo
from os2d_b200.structures import BoxList, FeatureMapSize
from os2d_b200.voc_eval import do_voc_evaluation

n_images, n_labels = 600, 200
data = voc_inputs(3, n_images=n_images, n_labels=n_labels)
preds, gts, arrays = [], [], ([], [], [], [], [], [])
for (pb, pl, ps, psize, gt, gl, gd, gsize) in data:
    b = BoxList(pb.cuda(), FeatureMapSize(w=psize[0], h=psize[1]))
    b.add_field("labels", pl.cuda())
    b.add_field("scores", ps.cuda())
    preds.append(b)
    t = BoxList(gt, FeatureMapSize(w=gsize[0], h=gsize[1]))
    t.add_field("labels", gl)
    t.add_field("difficult", gd)
    gts.append(t)
    rw, rh = float(gsize[0]) / psize[0], float(gsize[1]) / psize[1]
    scaled = pb * rw if rw == rh else pb * torch.tensor([rw, rh, rw, rh])
    for lst, v in zip(arrays, (scaled.numpy(), pl.numpy(), ps.numpy(), gt.numpy(), gl.numpy(), gd.numpy())):
        lst.append(v)
n_det = sum(len(x) for x in arrays[1])
do_voc_evaluation(preds, gts)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    r = do_voc_evaluation(preds, gts)
torch.cuda.synchronize()
t_dev = (time.perf_counter() - t0) / 5
t0 = time.perf_counter()
ro = vo.eval_detection_voc(*arrays)
t_cpu = time.perf_counter() - t0
print(json.dumps({"what": "do_voc_evaluation, {} images, {} labels, {} detections (wall clock, result dict on the host)".format(n_images, n_labels, n_det),
                  "device_ms": t_dev * 1e3, "cpu_oracle_ms": t_cpu * 1e3, "images_per_s_device": n_images / t_dev,
                  "images_per_s_cpu": n_images / t_cpu, "map_device": float(r["map"]), "map_cpu": float(ro["map"]),
                  "abs_diff_map": abs(float(r["map"]) - float(ro["map"]))}))
