"""Ring and RingRoot: drop-in mirrors of dot_ring/vrf/ring/members.py:18-81 and dot_ring/vrf/ring/root.py:13-112.

``Ring(keys, params)`` ingests the keys on the GPU (batched decompression + prime-subgroup check,
invalid keys -> padding point), builds the public vector ``PK || padding || 2^i*B || zeros`` and, in the
same device object, the three fixed columns (px, py, s): coefficients, KZG commitments (= the ring
root) and their 4x low-degree extensions.  The reference recomputes the latter on every proof
(constraints.py:43-62); here they live in HBM for the lifetime of the ring.
"""

from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass
from functools import lru_cache
from typing import Any

from . import _native
from .curve import point_to_string
from .engine import Engine, EnginePool, PooledRingNative, default_engine
from .params import RingProofParams


def _native_ring(engine: Engine, keys: Sequence[bytes], params: RingProofParams) -> _native.NativeRing:
    p = params.cv.curve.params
    aux = p.auxiliary_points
    return _native.NativeRing(
        engine.srs,
        [bytes(k) for k in keys],
        domain_size=params.domain_size,
        max_ring_size=params.max_ring_size,
        padding_rows=params.padding_rows,
        omega=params.omega,
        radix_omega=params.radix_omega,
        seed=aux.accumulator_base,
        blinding_base=aux.blinding_base,
        padding_point=aux.padding_point,
        generator=p.generator,
        suite_id=p.suite_id,
        h2c_dst=p.hash_to_curve.dst,
        hash_name=p.hash_name,
    )


class Ring:
    nm_points: tuple[tuple[int, int], ...]
    params: RingProofParams

    def __init__(self, keys: Sequence[bytes], params: RingProofParams | None = None, engine: "Engine | EnginePool | None" = None) -> None:
        if params is None:
            params = RingProofParams.from_ring_size(len(keys))
        self.params = params
        self.engine = engine or default_engine()
        aux = params.cv.curve.params.auxiliary_points
        if not aux.padding_point:
            raise ValueError("padding point is not configured in curve parameters")
        if len(keys) > params.max_ring_size:
            raise ValueError(f"ring size {len(keys)} exceeds max supported size {params.max_ring_size}")
        self.keys = tuple(bytes(k) for k in keys)
        # keys of the wrong length can never decode; the reference maps them to the padding point as well
        normalised = [k if len(k) == 32 else b"\xff" * 32 for k in self.keys]
        if isinstance(self.engine, EnginePool):  # one replica per GPU; batches are sharded over them (engine.py)
            self.native = PooledRingNative(self.engine, lambda eng: _native_ring(eng, normalised, params))
        else:
            self.native = _native_ring(self.engine, normalised, params)
        self.nm_points = tuple(self.native.points())
        self._index: dict[tuple[int, int], int] | None = None

    @classmethod
    def from_keys(cls, keys: Sequence[bytes], params: RingProofParams | None = None) -> "Ring":
        """Cached ring for a stable encoded key vector (members.py:57-59)."""
        return _ring(tuple(bytes(key) for key in keys))

    def _decode_key(self, key: bytes) -> tuple[int, int] | None:
        if len(key) != 32:
            return None
        pt = self.engine.ctx.te_decode([bytes(key)], checked=True)[0]
        if pt is None or pt == (0, 1):
            return None
        return pt

    def index_of(self, key: bytes) -> int:
        """members.py:71-81."""
        padding_point = self.params.cv.curve.params.auxiliary_points.padding_point
        point = self._decode_key(key)
        if point is None:
            raise ValueError("invalid ring key")
        if point == padding_point:
            raise ValueError("producer key is not in ring")
        if self._index is None:
            index: dict[tuple[int, int], int] = {}
            for i, pt in enumerate(self.nm_points[: self.params.max_ring_size]):
                index.setdefault(pt, i)
            self._index = index
        try:
            return self._index[point]
        except KeyError as exc:
            raise ValueError("producer key is not in ring") from exc


@lru_cache(maxsize=2)
def _params(keys_len: int) -> RingProofParams:
    return RingProofParams.from_ring_size(keys_len)


@lru_cache(maxsize=8)
def _ring(keys: tuple[bytes, ...]) -> Ring:
    return Ring(keys, _params(len(keys)))


@dataclass
class Column:
    """Commitment holder mirroring dot_ring/ring_proof/columns/columns.py:21-60 (device keeps evals/coeffs)."""

    name: str
    _commitment: bytes | None = None
    size: int = 512

    @property
    def commitment(self) -> bytes:
        if self._commitment is None:
            raise ValueError(f"{self.name} commitment is not set")
        return self._commitment


@dataclass
class RingRoot:
    px: Column
    py: Column
    s: Column
    params: RingProofParams | None = None

    @classmethod
    def from_ring(cls, ring: Ring, params: RingProofParams | None = None) -> "RingRoot":
        """root.py:21-44: the three commitments were produced when the ring was ingested."""
        if params is None:
            params = ring.params
        if (params.domain_size, params.max_ring_size) != (ring.params.domain_size, ring.params.max_ring_size):
            ring = Ring(ring.keys, params, ring.engine)
        raw = ring.native.fixed_commitments()
        n = params.domain_size
        return cls(
            px=Column("px", raw[0:96], n), py=Column("py", raw[96:192], n), s=Column("s", raw[192:288], n), params=params
        )

    def fixed_commitments(self) -> list[Any]:
        return [self.px.commitment, self.py.commitment, self.s.commitment]

    def verifier_transcript_prefix(self, transcript_challenge: bytes | None = None):
        """root.py:54-71: transcript state after absorbing the verifier key = [1]_1 | [1]_2 | [tau]_2 | C_px | C_py | C_s (uncompressed)."""
        if self.params is None:
            raise ValueError("Ring root verifier transcript requires ring proof parameters")
        from .transcript import FiatShamirTranscript

        srs = default_engine().srs_bytes
        if transcript_challenge is None:
            transcript_challenge = self.params.cv.curve.params.suite_id
        transcript = FiatShamirTranscript(self.params.prime, transcript_challenge)
        transcript.absorb_labeled(b"vk", srs.g1_be96[:96] + srs.g2_be192 + b"".join(self.fixed_commitments()))
        return transcript

    @staticmethod
    def encoded_len(params: RingProofParams | None = None) -> int:
        return 3 * 48

    def encode(self) -> bytes:
        from .kzg import KZG

        pcs = self.params.pcs if self.params is not None else KZG
        return pcs.compress_g1(self.px.commitment + self.py.commitment + self.s.commitment)

    @classmethod
    def decode(cls, data: bytes, ring: Ring | RingProofParams | None = None) -> "RingRoot":
        """root.py:89-109."""
        params = ring.params if isinstance(ring, Ring) else ring
        if params is None:
            params = RingProofParams()
        expected = cls.encoded_len(params)
        if len(data) != expected:
            raise ValueError(f"invalid ring root length: ring root must be exactly {expected} bytes, got {len(data)}")
        pts = params.pcs.decompress_g1_batch(bytes(data))
        n = params.domain_size
        return cls(px=Column("px", pts[0], n), py=Column("py", pts[1], n), s=Column("s", pts[2], n), params=params)

    def matches_ring(self, ring: Ring) -> bool:
        """root.py:111-112."""
        return RingRoot.from_ring(ring).encode() == self.encode()


__all__ = ["Ring", "RingRoot", "Column", "point_to_string"]
