"""Seeded synthetic TransformNet weights for the measurement scripts (same generator sequence as the test helper
oracle.head_oracle.random_transform_net, restated here so that nothing outside tests/ imports oracle/)."""
import math

import torch


def seeded_transform_net(out_dim, seed=0, spread=0.02):
    """Reference-shaped TransformNet state dict (keys conv.0/1/3/4, linear) with a non-identity output."""
    g = torch.Generator().manual_seed(seed)
    tn = {}

    def conv(name, co, ci, k):
        bound = 1.0 / math.sqrt(ci * k * k)
        tn[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        tn[name + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound

    def bn(name, c):
        tn[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        tn[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        tn[name + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
        tn[name + ".running_var"] = 0.01 + 0.05 * torch.rand(c, generator=g)

    conv("conv.0", 128, 225, 7)
    bn("conv.1", 128)
    conv("conv.3", 64, 128, 5)
    bn("conv.4", 64)
    tn["linear.weight"] = spread * torch.randn(out_dim, 64, 5, 5, generator=g)
    bias = torch.zeros(out_dim)
    bias[0] = 1
    bias[4 if out_dim == 6 else 2] = 1
    tn["linear.bias"] = bias
    return tn
