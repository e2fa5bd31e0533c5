"""CPU: the vectorised coefficient tables of the product (os2d_b200/pyramid.py) equal the scalar restatement of Pillow's
precompute_coeffs in the oracle, which is itself pinned against Pillow (tests/test_oracle_resize.py)."""
import numpy as np
import pytest

from oracle import resize_oracle as ro


@pytest.mark.parametrize("in_size,out_size", [(64, 32), (64, 96), (97, 41), (61, 77), (50, 50), (47, 160), (200, 37), (150, 29),
                                              (31, 7), (17, 5), (123, 124), (77, 76), (640, 224), (1280, 2048), (1333, 800),
                                              (5, 1), (1, 9), (3000, 1700)])
def test_coefficient_tables_equal_oracle(in_size, out_size):
    from os2d_b200.pyramid import resize_coefficients
    b, c, k = resize_coefficients(in_size, out_size)
    bo, co, ko = ro.bilinear_coeffs(in_size, out_size)
    assert k == ko and b.dtype == np.int32 and c.dtype == np.int32
    assert np.array_equal(b, bo) and np.array_equal(c, co)
    assert (c.sum(axis=1) > 0).all() and abs(int(c.sum(axis=1).max()) - (1 << 22)) <= c.shape[1]
