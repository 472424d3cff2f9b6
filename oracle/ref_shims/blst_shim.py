"""``dot_ring.blst`` stand-in: the SWIG API subset the reference calls,
backed by the oracle's BLS12-381 restatement.  Bad encodings raise RuntimeError like blst does
(kzg.py:141-144 maps that to ValueError)."""

from __future__ import annotations

from oracle import bls12_381 as B


class P1_Affine:
    def __init__(self, data=None):
        if data is None:
            self.pt = None
        elif isinstance(data, P1):
            self.pt = B.g1_to_affine(data.pt)
        else:
            try:
                self.pt = B.g1_to_affine(B.g1_decompress(bytes(data)))
            except ValueError as exc:
                raise RuntimeError("BLST_BAD_ENCODING") from exc

    def to_jacobian(self):
        return P1(self)

    def serialize(self):
        return B.g1_serialize(None if self.pt is None else (self.pt[0], self.pt[1], 1))

    def compress(self):
        return B.g1_compress(None if self.pt is None else (self.pt[0], self.pt[1], 1))


class P1:
    def __init__(self, src=None):
        if src is None:
            self.pt = None
        elif isinstance(src, P1_Affine):
            self.pt = None if src.pt is None else (src.pt[0], src.pt[1], 1)
        elif isinstance(src, P1):
            self.pt = src.pt
        else:
            self.pt = P1(P1_Affine(src)).pt

    def dup(self):
        return P1(self)

    def add(self, other):
        if isinstance(other, P1_Affine):
            other = P1(other)
        self.pt = B.g1_add(self.pt, other.pt)
        return self

    def neg(self):
        self.pt = B.g1_neg(self.pt)
        return self

    def mult(self, scalar):
        self.pt = B.g1_mul(self.pt, int(scalar))
        return self

    def to_affine(self):
        return P1_Affine(self)

    def serialize(self):
        return B.g1_serialize(self.pt)

    def compress(self):
        return B.g1_compress(self.pt)

    def is_inf(self):
        return self.pt is None


class _Memory:
    def __init__(self, pts):
        self.pts = pts

    def __getitem__(self, idx):
        if not isinstance(idx, slice):
            raise TypeError("memory supports slicing only")
        return _Memory(self.pts[idx])

    def __len__(self):
        return len(self.pts)


class P1_Affines:
    @staticmethod
    def as_memory(points):
        return _Memory([B.g1_to_affine(p.pt) if isinstance(p, P1) else p.pt for p in points])

    @staticmethod
    def mult_pippenger(memory, scalars):
        out = P1()
        out.pt = B.g1_msm(memory.pts, [int(s) for s in scalars])
        return out


class P2_Affine:
    def __init__(self, data=None):
        if isinstance(data, P2):
            self.pt = data.pt
        elif data is None:
            self.pt = None
        else:
            data = bytes(data)
            try:
                self.pt = B.g2_decompress(data) if len(data) == 96 else B.g2_from_uncompressed(data)
            except ValueError as exc:
                raise RuntimeError("BLST_BAD_ENCODING") from exc


class P2:
    def __init__(self, src=None):
        if src is None:
            self.pt = None
        elif isinstance(src, (P2, P2_Affine)):
            self.pt = src.pt
        else:
            self.pt = P2_Affine(src).pt

    def dup(self):
        return P2(self)

    def add(self, other):
        self.pt = B.g2_add(self.pt, other.pt)
        return self

    def neg(self):
        self.pt = B.g2_neg(self.pt)
        return self

    def mult(self, scalar):
        self.pt = B.g2_mul(self.pt, int(scalar))
        return self

    def to_affine(self):
        return P2_Affine(self)

    def serialize(self):
        return B.g2_serialize(self.pt)


class PT:
    def __init__(self, q: P2_Affine, p: P1_Affine):
        self.f = B.miller_loop(q.pt, p.pt)

    @staticmethod
    def finalverify(a, b):
        return B.final_verify(a.f, b.f)
