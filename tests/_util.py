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
