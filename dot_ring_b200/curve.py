"""Bandersnatch suite constants and the CurveVariant facade (host side).

Mirrors the parts of dot_ring/curve/specs/bandersnatch.py:48-107,289-293 and
dot_ring/curve/curve.py:359-399 that the ring-proof path touches.  All group arithmetic is done
by the CUDA library; this module only carries constants, hashing glue and byte codecs.
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass

FIELD_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SUBGROUP_ORDER = 0x1CFB69D4CA675F520CCE760202687600FF8F87007419047174FD06B52876E7E1


@dataclass(frozen=True)
class AuxiliaryPoints:
    blinding_base: tuple[int, int]
    accumulator_base: tuple[int, int]
    padding_point: tuple[int, int]


@dataclass(frozen=True)
class HashToCurve:
    dst: bytes


@dataclass(frozen=True)
class Encoding:
    point_len: int = 32
    challenge_len: int = 16


@dataclass(frozen=True)
class BandersnatchParams:
    field_modulus: int
    subgroup_order: int
    cofactor: int
    suite_id: bytes
    generator: tuple[int, int]
    a: int
    d: int
    hash_to_curve: HashToCurve
    encoding: Encoding
    auxiliary_points: AuxiliaryPoints
    hash_name: str = "sha512"  # "sha512" | "shake128" (specs/bandersnatch.py:48-144)

    @property
    def hash_fn(self):
        return hashlib.sha512 if self.hash_name == "sha512" else hashlib.shake_128


@dataclass(frozen=True)
class _CurveHolder:
    params: BandersnatchParams


class CurveVariant:
    """``Bandersnatch``-like suite handle: ``.name``, ``.curve.params``, key helpers."""

    def __init__(self, name: str, params: BandersnatchParams):
        self.name = name
        self.curve = _CurveHolder(params)

    def __repr__(self) -> str:
        return f"CurveVariant({self.name})"

    # dot_ring/curve/curve.py:377-384
    def public_key_from_secret(self, secret_key: bytes) -> bytes:
        if not isinstance(secret_key, (bytes, bytearray)):
            raise TypeError("secret_key must be bytes")
        return self.public_keys_from_secrets([bytes(secret_key)])[0]

    def public_keys_from_secrets(self, secret_keys: list[bytes]) -> list[bytes]:
        from .engine import default_engine

        gen = point_to_string(self.curve.params.generator)
        out = default_engine().ctx.te_mul([gen], [int.from_bytes(sk, "little") for sk in secret_keys])
        return [bytes(o) for o in out]

    def msm(self, points: list[bytes], scalars: list[int]) -> bytes:
        """`BandersnatchPoint.msm` (specs/bandersnatch.py:194-286) on 32-byte encodings: sum_i k_i * P_i on the device."""
        from .engine import default_engine

        return default_engine().ctx.te_msm([bytes(p) for p in points], [int(k) for k in scalars])

    # dot_ring/curve/curve.py:386-399 + dot_ring/vrf/primitives.py:147-162
    def secret_from_seed(self, seed: bytes) -> tuple[bytes, bytes]:
        if not isinstance(seed, (bytes, bytearray)):
            raise TypeError("seed must be bytes")
        from .transcript import secret_scalar_from_seed

        secret = secret_scalar_from_seed(self, bytes(seed))
        secret_key = secret.to_bytes(32, "little")
        return self.public_key_from_secret(secret_key), secret_key


def point_to_string(pt: tuple[int, int]) -> bytes:
    """dot_ring/curve/point.py:150-176: y little-endian, x-sign in bit 7 of the last byte."""
    x, y = pt
    out = bytearray(int(y).to_bytes(32, "little"))
    if x > (-x) % FIELD_MODULUS:
        out[31] |= 0x80
    return bytes(out)


BANDERSNATCH_PARAMS = BandersnatchParams(
    field_modulus=FIELD_MODULUS,
    subgroup_order=SUBGROUP_ORDER,
    cofactor=4,
    suite_id=b"Bandersnatch-SHA512-ELL2-v1",
    generator=(
        18886178867200960497001835917649091219057080094937609519140440539760939937304,
        19188667384257783945677642223292697773471335439753913231509108946878080696678,
    ),
    a=-5,
    d=0x6389C12633C267CBC66E3BF86BE3B6D8CB66677177E54F92B369F2F5188D58E7,
    hash_to_curve=HashToCurve(dst=b"Bandersnatch-SHA512-ELL2-v1\x60"),
    encoding=Encoding(),
    auxiliary_points=AuxiliaryPoints(
        blinding_base=(
            23335687741101763108036518445642207119627658113885888016488710494487028845889,
            5552214580375038693022409684979828600325210968745774080859660443337357929963,
        ),
        accumulator_base=(
            14056632001415368875257708737821299882600475929746323097150942355715730684350,
            10322661992765989500407719465917595459409463902187386706652408883505670839210,
        ),
        padding_point=(
            26913883415342152801331916189968962157924271221160514298872262294143390094043,
            30874728313203001508631936119690348239461579770372782660098261717479009115354,
        ),
    ),
)

Bandersnatch = CurveVariant("Bandersnatch", BANDERSNATCH_PARAMS)

# specs/bandersnatch.py:108-144: same curve and auxiliary points, SHAKE128 transcripts and XOF hash-to-curve
import dataclasses as _dc

BANDERSNATCH_SHAKE128_PARAMS = _dc.replace(
    BANDERSNATCH_PARAMS,
    suite_id=b"Bandersnatch-SHAKE128-ELL2-v1",
    hash_to_curve=HashToCurve(dst=b"Bandersnatch-SHAKE128-ELL2-v1\x60"),
    hash_name="shake128",
    auxiliary_points=AuxiliaryPoints(  # derived with the suite's own hash-to-curve (specs/bandersnatch.py:128-139)
        blinding_base=(
            6153734995852631824944342602386415873379775188383988340041079006556670120775,
            27204351599954061630605768787803524395123895650061061132592995395630473050754,
        ),
        accumulator_base=(
            27631238720955528589004064829276283990465032040945349648037876197995278250917,
            37605358688136619817560700742505556266961225274493904038881144193539047100140,
        ),
        padding_point=(
            1834402953989431481748983728202937234471322740714585873803966488035889514523,
            52100941849053769665273763352270294131006971127418863694682093199651869272752,
        ),
    ),
)
Bandersnatch_SHAKE128 = CurveVariant("Bandersnatch_SHAKE128", BANDERSNATCH_SHAKE128_PARAMS)
