"""A tiny synthetic dataset in the GroZi-3.2k layout the reference reads (os2d/data/dataset.py:42, 76-123): CSV with the columns
imageid, imagefilename, classid, classfilename, gtbboxid, difficult, lx, ty, rx, by, split under <data>/grozi/classes/grozi.csv,
class images under <data>/grozi/classes/images/, scene images (longest side 3264, as on disk in the real dataset, so that the
reference does not resize them - its resize path uses the removed Image.ANTIALIAS, dataset.py:671) under <data>/grozi/src/3264/.
Scenes are pastes of the class images on textured backgrounds, so the ground truth is exact."""
import os

import numpy as np
from PIL import Image


def make(data_root, n_images=2, n_classes=3, seed=0):
    rng = np.random.RandomState(seed)
    base = os.path.join(data_root, "grozi")
    os.makedirs(os.path.join(base, "classes", "images"), exist_ok=True)
    os.makedirs(os.path.join(base, "src", "3264"), exist_ok=True)
    class_imgs = []
    for c in range(n_classes):
        h, w = 300 + 40 * c, 260 - 30 * c
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([(np.sin(xx / (7.0 + c)) * 90 + 128), (np.cos(yy / (5.0 + 2 * c)) * 90 + 128),
                        ((xx + yy) % (20 + 5 * c)) * (255.0 / (20 + 5 * c))], axis=2).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(base, "classes", "images", "class_%d.jpg" % c), quality=95)
        class_imgs.append(img)
    rows = ["imageid,imagefilename,classid,classfilename,gtbboxid,difficult,lx,ty,rx,by,split"]
    box_id = 0
    W, H = 3264, 2448
    for i in range(n_images):
        scene = (rng.rand(H // 16, W // 16, 3) * 60 + 90).astype(np.uint8)
        scene = np.asarray(Image.fromarray(scene).resize((W, H), Image.BILINEAR)).copy()
        for c in range(n_classes):
            if (i + c) % n_classes == n_classes - 1 and n_classes > 1:
                continue                                          # not every class in every image
            ch, cw = class_imgs[c].shape[:2]
            s = 2.0 + 0.5 * ((i + c) % 3)                           # objects ~ 2-3x the class image: ~600-900 px in a 3264 px scene
            ph, pw = int(ch * s), int(cw * s)
            y0, x0 = int(rng.randint(0, H - ph)), int(rng.randint(0, W - pw))
            patch = np.asarray(Image.fromarray(class_imgs[c]).resize((pw, ph), Image.BILINEAR))
            scene[y0:y0 + ph, x0:x0 + pw] = patch
            rows.append("%d,scene_%d.jpg,%d,class_%d.jpg,%d,0,%.6f,%.6f,%.6f,%.6f,val-new-cl" % (
                i, i, c, c, box_id, x0 / W, y0 / H, (x0 + pw) / W, (y0 + ph) / H))
            box_id += 1
        Image.fromarray(scene).save(os.path.join(base, "src", "3264", "scene_%d.jpg" % i), quality=90)
    with open(os.path.join(base, "classes", "grozi.csv"), "w") as f:
        f.write("\n".join(rows) + "\n")
    return base
