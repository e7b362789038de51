"""Import-only stand-in for the `galois` package (absent from this image, no network).

Used ONLY by tests/golden/make_golden.py so that the read-only reference at /root/reference can be
imported in this container to generate structural golden fixtures.  Nothing here does field
arithmetic: any reference path that needs real GF(2) algebra (logical operators) raises.
"""
import numpy as np


class FieldArray(np.ndarray):
    def __new__(cls, a):
        raise NotImplementedError("galois stub: GF(2) arithmetic is not available in this container")


def GF(order):
    assert order == 2
    return FieldArray


GF2 = FieldArray


class Poly:  # pragma: no cover - only imported by generators we do not call
    def __init__(self, *a, **k):
        raise NotImplementedError("galois stub")
