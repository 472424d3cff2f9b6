"""KZG over the device-resident SRS: the `PCS` seam of the reference (dot_ring/ring_proof/pcs/protocol.py:10-40).

G1 points are carried as 96-byte zcash-uncompressed ``bytes`` (infinity = 0x40 || 0...), which is what the
reference absorbs into its transcript and what the C ABI speaks.  ``commit`` replaces
``blst.P1_Affines.mult_pippenger(srs.blst_g1_memory[:n], coeffs)`` (kzg.py:152-175).
"""

from __future__ import annotations

import hashlib
import secrets
from dataclasses import dataclass
from typing import NamedTuple

from .engine import default_engine

G1_INFINITY = bytes([0x40]) + bytes(95)


@dataclass(slots=True, frozen=True)
class Opening:
    """pcs/opening.py: the quotient commitment and the opened value."""

    proof: bytes
    y: int


class PcsVerification(NamedTuple):
    """pcs/utils.py:13-17."""

    commitment: bytes
    proof: bytes
    point: int
    value: int


class LinearPcsVerification(NamedTuple):
    """pcs/utils.py:20-24: the commitment is sum_i scalar_i * point_i."""

    commitment_terms: tuple
    proof: bytes
    point: int
    value: int


def _random_nonzero_coefficients(count: int, order: int) -> list[int]:
    """pcs/kzg.py:84-108: coefficient 0 is 1, the others are rejection-sampled from SHAKE256(32 random bytes | counter)."""
    coeffs = [1] if count > 0 else []
    width = (order.bit_length() + 7) // 8
    limit = (1 << (8 * width)) // order * order
    seed, counter = secrets.token_bytes(32), 0
    while len(coeffs) < count:
        raw = hashlib.shake_256(seed + counter.to_bytes(8, "little")).digest(2 * width * (count - len(coeffs)))
        counter += 1
        for off in range(0, len(raw), width):
            v = int.from_bytes(raw[off : off + width], "big")
            if v < limit and v % order and len(coeffs) < count:
                coeffs.append(v % order)
    return coeffs


class _Side:
    """sum_i k_i * P_i with equal points merged (pcs/kzg.py:27-53 merges by object identity; bytes compare by value)."""

    def __init__(self, order: int):
        self.order, self.terms = order, {}

    def add(self, point: bytes, scalar: int) -> None:
        scalar %= self.order
        if scalar:
            point = bytes(point)
            self.terms[point] = (self.terms.get(point, 0) + scalar) % self.order

    def vectors(self) -> tuple[list[bytes], list[int]]:
        items = [(p, k) for p, k in self.terms.items() if k]
        return [p for p, _ in items], [k for _, k in items]


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
    def open(cls, coeffs, x: int) -> Opening:
        """kzg.py:178-191: value at x and the commitment of the quotient by (X - x)."""
        return cls.open_batch([list(coeffs)], [x])[0]

    @classmethod
    def open_batch(cls, coeff_vectors, points) -> list[Opening]:
        eng = default_engine()
        vecs = [list(v) for v in coeff_vectors]
        if any(len(v) - 1 > eng.srs.size for v in vecs):
            raise ValueError("polynomial degree exceeds SRS size")
        return [Opening(p, y) for p, y in eng.srs.open(vecs, [int(x) for x in points])]

    @classmethod
    def _check(cls, lhs: _Side, rhs: _Side) -> bool:
        lp, ls = lhs.vectors()
        rp, rs = rhs.vectors()
        return default_engine().srs.pairing_check(lp, ls, rp, rs)

    @classmethod
    def _g1_generator(cls) -> bytes:
        return default_engine().srs_bytes.g1_be96[:96]

    @classmethod
    def verify(cls, commitment: bytes, proof: bytes, point: int, value: int) -> bool:
        """kzg.py:194-229: e(C - v [1]_1, [1]_2) == e(pi, [tau]_2 - z [1]_2), checked in the equivalent form
        e(C - v [1]_1 + z pi, [1]_2) == e(pi, [tau]_2) that the reference's own batch path uses (kzg.py:265-301)."""
        return cls._batch([PcsVerification(commitment, proof, point, value)], [1])

    @classmethod
    def _batch(cls, verifications, coeffs) -> bool:
        order = cls.scalar_modulus
        lhs, rhs = _Side(order), _Side(order)
        sum_v = 0
        for coeff, (commitment, proof, point, value) in zip(coeffs, verifications):
            lhs.add(commitment, coeff)
            sum_v = (sum_v + coeff * value) % order
            lhs.add(proof, coeff * point)
            rhs.add(proof, coeff)
        lhs.add(cls._g1_generator(), -sum_v)
        return cls._check(lhs, rhs)

    @classmethod
    def batch_verify(cls, verifications) -> bool:
        """kzg.py:231-302: random linear combination of the openings, one pairing equation."""
        verifications = list(verifications)
        if not verifications:
            return True
        if len(verifications) == 1:
            return cls.verify(*verifications[0])
        return cls._batch(verifications, _random_nonzero_coefficients(len(verifications), cls.scalar_modulus))

    @classmethod
    def batch_verify_linear_preconverted(cls, verifications) -> bool:
        """kzg.py:304-338 (+ `_aggregate_linear_batch`, :56-81): openings of commitments given as linear combinations."""
        verifications = list(verifications)
        if not verifications:
            return True
        order = cls.scalar_modulus
        coeffs = _random_nonzero_coefficients(len(verifications), order)
        lhs, rhs = _Side(order), _Side(order)
        sum_v = 0
        for coeff, ver in zip(coeffs, verifications):
            for commitment, scalar in ver.commitment_terms:
                lhs.add(commitment, coeff * scalar)
            sum_v = (sum_v + coeff * ver.value) % order
            lhs.add(ver.proof, coeff * ver.point)
            rhs.add(ver.proof, coeff)
        lhs.add(cls._g1_generator(), -sum_v)
        return cls._check(lhs, rhs)

    @classmethod
    def msm_g1(cls, points, scalars) -> bytes:
        """kzg.py:147-149 over arbitrary points (96-byte uncompressed each): device bucket-method MSM."""
        pts = list(points)
        if len(pts) != len(scalars):
            raise ValueError("points and scalars must have the same length")
        eng = default_engine()
        data, ks = b"".join(bytes(p) for p in pts), [int(k) for k in scalars]
        if hasattr(eng, "g1_msm"):  # EnginePool: very large point sets are split by range over the devices
            return eng.g1_msm(data, ks)
        return eng.ctx.g1_msm(data, ks)

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
