"""Shared helpers of the test-suite (synthetic inputs, tolerances, golden loading)."""
import os

import numpy as np
import torch

from oracle import head_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = "/root/reference"

# Parity bar of BASELINE.json north_star: 1e-3 relative fp32 tolerance.  Metric: per
# output tensor max|a-b| <= TOL * max|b|.
TOL = 1e-3

VARIANTS = [("affine_inverse", False, True), ("affine", False, False), ("simple", True, False),
            ("simple_inverse", True, True)]


def rel_to_max(a, b):
    a = torch.as_tensor(a).float()
    b = torch.as_tensor(b).float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def load_head_golden(name):
    z = np.load(os.path.join(GOLDEN, "head_%s.npz" % name))
    d = {k: z[k] for k in z.files}
    cms = []
    i = 0
    while "class_map_%d" % i in d:
        cms.append(torch.from_numpy(d["class_map_%d" % i]))
        i += 1
    P = 4 if int(d["simple"]) else 6
    tn = ho.random_transform_net(P, seed=int(d["tn_seed"]), spread=float(d["tn_spread"]))
    chk = float(sum(v.double().abs().sum() for v in tn.values()))
    assert abs(chk - float(d["tn_checksum"])) <= 1e-9 * abs(chk), "seeded TransformNet weights differ from the golden run"
    return d, cms, torch.from_numpy(d["fm"]), tn


def synth_inputs(seed, B, H, W, sizes, D=1024):
    g = torch.Generator().manual_seed(seed)
    cms = [(torch.randn(1, D, h, w, generator=g) * 0.5 + 0.2).relu() for (h, w) in sizes]
    fm = (torch.randn(B, D, H, W, generator=g) * 0.5 + 0.2).relu()
    return cms, fm


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "os2d"))


def singular_params(seed, NB, P, H, W, singular_at):
    """Same seeded parameters as tests/golden/make_golden.py::singular_params (exactly singular matrices planted)."""
    g = torch.Generator().manual_seed(seed)
    ident = torch.tensor([1., 0, 0, 0, 1, 0] if P == 6 else [1., 0, 1, 0]).view(1, P, 1, 1)
    p = ident + 0.2 * torch.randn(NB, P, H, W, generator=g)
    kinds6 = [[0., 0, 0.3, 0, 0, -0.2], [1., 2, 0.1, 1, 2, 0.5], [0.5, 0.25, 0.1, 2, 1, 0.5]]
    kinds4 = [[0., 0.3, 1.1, -0.2], [0.9, 0.1, 0., 0.5], [0., 0., 0., 0.]]
    for i, (n, y, x) in enumerate(singular_at):
        p[n, :, y, x] = torch.tensor((kinds6 if P == 6 else kinds4)[i % 3])
    return p


def voc_inputs(seed, n_images=14, n_labels=9, quant=0):
    """Same generator as tests/golden/make_golden.py::voc_inputs.  Seeded detections / ground truth per image: (pred boxes, labels, scores, pred image size, gt boxes, labels,
    difficult, gt image size).  Label 3 never has ground truth, label 6 is never detected, image 2 has no detections,
    image 5 no ground truth; every third image predicts at another resolution (exercises BoxList.resize).
    quant > 0 quantises the scores (ties)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_images):
        W, H = 640 + 32 * (i % 3), 480 + 16 * (i % 4)
        ng = 0 if i == 5 else int(torch.randint(1, 9, (1,), generator=g))
        gxy = torch.rand(ng, 2, generator=g) * torch.tensor([W - 160.0, H - 120.0])
        gwh = 30 + torch.rand(ng, 2, generator=g) * 120
        gt = torch.cat([gxy, gxy + gwh], 1)
        gl = torch.randint(0, n_labels, (ng,), generator=g)
        gl[gl == 3] = 4
        gd = (torch.rand(ng, generator=g) < 0.25).to(torch.int64)
        nd = 0 if i == 2 else int(torch.randint(5, 60, (1,), generator=g))
        # half of the detections are jittered copies of ground-truth boxes (several per box: duplicates), half random
        src = torch.randint(0, max(ng, 1), (nd,), generator=g)
        jit = (torch.rand(nd, 4, generator=g) - 0.5) * 40
        near = (gt[src] + jit) if ng > 0 else torch.zeros(nd, 4)
        rxy = torch.rand(nd, 2, generator=g) * torch.tensor([W - 100.0, H - 100.0])
        rnd = torch.cat([rxy, rxy + 20 + torch.rand(nd, 2, generator=g) * 150], 1)
        use_near = (torch.rand(nd, generator=g) < 0.6) & (ng > 0)
        pb = torch.where(use_near[:, None], near, rnd)
        pb = torch.stack([torch.minimum(pb[:, 0], pb[:, 2]), torch.minimum(pb[:, 1], pb[:, 3]),
                          torch.maximum(pb[:, 0], pb[:, 2]), torch.maximum(pb[:, 1], pb[:, 3])], 1)
        pl = torch.where(use_near & (torch.rand(nd, generator=g) < 0.8), gl[src] if ng > 0 else torch.zeros(nd, dtype=torch.int64),
                         torch.randint(0, n_labels, (nd,), generator=g))
        pl[pl == 6] = 7
        ps = torch.rand(nd, generator=g)
        if quant:
            ps = (ps * quant).floor() / quant
        sx = 0.5 if i % 3 == 1 else 1.0             # predictions made on a smaller image: equal ratios (one multiply in
        sy = (0.5 if i % 2 else 0.25) if i % 3 == 1 else 1.0   # BoxList.resize) or different ratios (per-axis branch)
        psize = (int(W * sx), int(H * sy))
        out.append((pb * torch.tensor([sx, sy, sx, sy]), pl, ps, psize, gt, gl, gd, (W, H)))
    return out
