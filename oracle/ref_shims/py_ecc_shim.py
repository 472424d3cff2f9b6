"""py_ecc stand-in (API subset used at pcs/utils.py:3-58, pcs/srs.py:10,117-130, pcs/kzg.py:7,113,126)."""

from __future__ import annotations

import types

from oracle import bls12_381 as B


class FQ(int):
    def __new__(cls, v=0):
        return super().__new__(cls, int(v) % B.P)

    @property
    def n(self):
        return int(self)

    @classmethod
    def one(cls):
        return cls(1)

    @classmethod
    def zero(cls):
        return cls(0)


class FQ2:
    def __init__(self, coeffs):
        self.coeffs = tuple(FQ(c) for c in coeffs)

    def __eq__(self, other):
        return isinstance(other, FQ2) and self.coeffs == other.coeffs

    def __hash__(self):
        return hash(self.coeffs)


def normalize(pt):
    x, y, z = pt
    if isinstance(x, FQ2):
        if z == FQ2([1, 0]):
            return x, y
        raise NotImplementedError("projective G2 normalisation is not on the reference's path")
    a = B.g1_to_affine((int(x), int(y), int(z)))
    return FQ(a[0]), FQ(a[1])


def compress_G1(pt):
    x, y, z = pt
    if int(z) == 0:
        return (1 << 383) + (1 << 382)
    ax, ay = B.g1_to_affine((int(x), int(y), int(z)))
    return ax + ((ay * 2) // B.P) * (1 << 381) + (1 << 383)


def compress_G2(pt):
    x, y, z = pt
    x_re, x_im = (int(c) for c in x.coeffs)
    y_re, y_im = (int(c) for c in y.coeffs)
    a_flag = (y_im * 2) // B.P if y_im > 0 else (y_re * 2) // B.P
    return x_im + a_flag * (1 << 381) + (1 << 383), x_re


def register(modules) -> None:
    root = types.ModuleType("py_ecc")
    root.__path__ = []
    opt = types.ModuleType("py_ecc.optimized_bls12_381")
    opt.curve_order = B.R
    opt.field_modulus = B.P
    opt.FQ, opt.FQ2, opt.normalize = FQ, FQ2, normalize
    bls = types.ModuleType("py_ecc.bls")
    bls.__path__ = []
    pc = types.ModuleType("py_ecc.bls.point_compression")
    pc.compress_G1, pc.compress_G2 = compress_G1, compress_G2
    bls.point_compression = pc
    root.optimized_bls12_381, root.bls = opt, bls
    modules["py_ecc"] = root
    modules["py_ecc.optimized_bls12_381"] = opt
    modules["py_ecc.bls"] = bls
    modules["py_ecc.bls.point_compression"] = pc
