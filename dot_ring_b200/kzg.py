"""KZG over the device-resident SRS: the `PCS` seam of the reference (dot_ring/ring_proof/pcs/protocol.py:10-40).

G1 points are carried as 96-byte zcash-uncompressed ``bytes`` (infinity = 0x40 || 0...), which is what the
reference absorbs into its transcript and what the C ABI speaks.  ``commit`` replaces
``blst.P1_Affines.mult_pippenger(srs.blst_g1_memory[:n], coeffs)`` (kzg.py:152-175).
"""

from __future__ import annotations

from .engine import default_engine

G1_INFINITY = bytes([0x40]) + bytes(95)


class KZG:
    commitment_size = 48
    scalar_modulus = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

    @classmethod
    def commit(cls, coeffs) -> bytes:
        """kzg.py:152-175; coefficients may be unreduced ints (ops.py:215-220)."""
        return cls.commit_batch([list(coeffs)])[0]

    @classmethod
    def commit_batch(cls, coeff_vectors) -> list[bytes]:
        eng = default_engine()
        vecs = [list(v) for v in coeff_vectors]
        if any(len(v) > eng.srs.size for v in vecs):
            raise ValueError("polynomial degree exceeds SRS size")
        return eng.srs.commit(vecs)

    @classmethod
    def msm_g1(cls, points, scalars) -> bytes:
        """kzg.py:147-149 over arbitrary points (96-byte uncompressed each): device bucket-method MSM."""
        pts = list(points)
        if len(pts) != len(scalars):
            raise ValueError("points and scalars must have the same length")
        return default_engine().ctx.g1_msm(b"".join(bytes(p) for p in pts), [int(k) for k in scalars])

    @classmethod
    def compress_g1(cls, point: bytes) -> bytes:
        """kzg.py:129-131."""
        return default_engine().ctx.g1_compress(bytes(point))

    @classmethod
    def serialize_g1_uncompressed(cls, point: bytes) -> bytes:
        """kzg.py:133-135."""
        if len(point) != 96:
            raise ValueError("expected a 96-byte uncompressed G1 point")
        return bytes(point)

    @classmethod
    def decompress_g1(cls, data: bytes) -> bytes:
        """kzg.py:137-144."""
        if len(data) != cls.commitment_size:
            raise ValueError(f"invalid BLS12-381 G1 length: expected {cls.commitment_size}, got {len(data)}")
        out, ok = default_engine().ctx.g1_decompress(bytes(data))
        if not ok[0]:
            raise ValueError("invalid BLS12-381 G1 encoding")
        return out

    @classmethod
    def decompress_g1_batch(cls, data: bytes) -> list[bytes]:
        out, ok = default_engine().ctx.g1_decompress(bytes(data))
        if not all(ok):
            raise ValueError("invalid BLS12-381 G1 encoding")
        return [out[96 * i : 96 * i + 96] for i in range(len(ok))]

    @classmethod
    def normalize_g1(cls, point: bytes) -> tuple[int, int]:
        """kzg.py:121-127."""
        return int.from_bytes(point[:48], "big"), int.from_bytes(point[48:], "big")
