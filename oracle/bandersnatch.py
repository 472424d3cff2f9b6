"""Bandersnatch twisted-Edwards group, codec and Elligator2 (oracle; test infrastructure only).

Restates, on Python ints:
  dot_ring/curve/specs/bandersnatch.py:48-144        (suite constants, both hash suites)
  dot_ring/curve/twisted_edwards/te_affine_point.py:69-316  (group law, clear cofactor, from_mont, x-recover)
  dot_ring/curve/twisted_edwards/te_curve.py:48-95   (Elligator 2 map)
  dot_ring/curve/curve.py:56-67,110-185              (valid_point, hash_to_field, expand_message_xmd/xof)
  dot_ring/curve/point.py:150-214                    (point <-> 32-byte string)
  dot_ring/curve/native_field/bandersnatch_te.pyx:127-174,421-477  (extended add/double, Tonelli-Shanks)
Group results are unique, so the scalar-multiplication schedule (GLV, windows) is not restated.
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # base field = BLS12-381 Fr
N = 0x1CFB69D4CA675F520CCE760202687600FF8F87007419047174FD06B52876E7E1  # prime subgroup order
COFACTOR = 4
A = -5 % P
D = 0x6389C12633C267CBC66E3BF86BE3B6D8CB66677177E54F92B369F2F5188D58E7
GENERATOR = (
    18886178867200960497001835917649091219057080094937609519140440539760939937304,
    19188667384257783945677642223292697773471335439753913231509108946878080696678,
)
IDENTITY = (0, 1)


@dataclass(frozen=True)
class Suite:
    """Per-hash-suite constants (specs/bandersnatch.py:66-144)."""

    name: str
    suite_id: bytes
    hash_name: str  # "sha512" | "shake128"
    dst: bytes
    blinding_base: tuple[int, int]
    accumulator_base: tuple[int, int]
    padding_point: tuple[int, int]

    def new_hash(self):
        return hashlib.sha512() if self.hash_name == "sha512" else hashlib.shake_128()


SHA512 = Suite(
    name="Bandersnatch",
    suite_id=b"Bandersnatch-SHA512-ELL2-v1",
    hash_name="sha512",
    dst=b"Bandersnatch-SHA512-ELL2-v1\x60",
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
)

SHAKE128 = Suite(
    name="Bandersnatch_SHAKE128",
    suite_id=b"Bandersnatch-SHAKE128-ELL2-v1",
    hash_name="shake128",
    dst=b"Bandersnatch-SHAKE128-ELL2-v1\x60",
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
)

# ---- field helpers ----------------------------------------------------------


def fr_is_square(v: int) -> bool:
    v %= P
    return v == 0 or pow(v, (P - 1) // 2, P) == 1


def fr_sqrt(v: int):
    """Any square root in Fr, or None (bandersnatch_te.pyx:421-477; callers normalise the sign)."""
    v %= P
    if v == 0:
        return 0
    if pow(v, (P - 1) // 2, P) != 1:
        return None
    q, s = P - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 5  # non-residue used by the reference's native routine
    c = pow(z, q, P)
    x = pow(v, (q + 1) // 2, P)
    t = pow(v, q, P)
    m = s
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P
            i += 1
        b = pow(c, 1 << (m - i - 1), P)
        x = x * b % P
        c = b * b % P
        t = t * c % P
        m = i
    return x


# ---- group law --------------------------------------------------------------


def is_on_curve(pt) -> bool:
    x, y = pt
    x2, y2 = x * x % P, y * y % P
    return (A * x2 + y2 - 1 - D * x2 % P * y2) % P == 0


def add(p1, p2):
    """Affine addition (te_affine_point.py:69-114); complete on the prime-order subgroup."""
    x1, y1 = p1
    x2, y2 = p2
    x1x2 = x1 * x2 % P
    y1y2 = y1 * y2 % P
    t = D * x1x2 % P * y1y2 % P
    x3 = (x1 * y2 + x2 * y1) * pow(1 + t, -1, P) % P
    y3 = (y1y2 - A * x1x2) * pow(1 - t, -1, P) % P
    return (x3, y3)


def neg(pt):
    return ((-pt[0]) % P, pt[1])


def _ext_add(p1, p2):
    # add-2008-hwcd on extended coordinates (X, Y, Z, T) for general a
    X1, Y1, Z1, T1 = p1
    X2, Y2, Z2, T2 = p2
    Aa = X1 * X2 % P
    Bb = Y1 * Y2 % P
    C = T1 * D % P * T2 % P
    Dd = Z1 * Z2 % P
    E = ((X1 + Y1) * (X2 + Y2) - Aa - Bb) % P
    F = (Dd - C) % P
    G = (Dd + C) % P
    H = (Bb - A * Aa) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def _ext_double(p1):
    X1, Y1, Z1, _ = p1
    Aa = X1 * X1 % P
    Bb = Y1 * Y1 % P
    C = 2 * Z1 * Z1 % P
    Dd = A * Aa % P
    E = ((X1 + Y1) * (X1 + Y1) - Aa - Bb) % P
    G = (Dd + Bb) % P
    F = (G - C) % P
    H = (Dd - Bb) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def _to_ext(pt):
    return (pt[0], pt[1], 1, pt[0] * pt[1] % P)


def _from_ext(e):
    zi = pow(e[2], -1, P)
    return (e[0] * zi % P, e[1] * zi % P)


def mul_raw(pt, k: int):
    """k*pt for a non-negative integer k (no reduction)."""
    acc = (0, 1, 1, 0)
    base = _to_ext(pt)
    while k:
        if k & 1:
            acc = _ext_add(acc, base)
        base = _ext_double(base)
        k >>= 1
    return _from_ext(acc)


def mul(pt, k: int):
    """Scalar multiple with k reduced mod the subgroup order (specs/bandersnatch.py:177-191)."""
    return mul_raw(pt, k % N)


def msm(points, scalars):
    """sum k_i * P_i (specs/bandersnatch.py:193-286; schedule not restated)."""
    acc = (0, 1, 1, 0)
    for pt, k in zip(points, scalars, strict=True):
        k %= N
        if k == 0:
            continue
        base = _to_ext(pt)
        part = (0, 1, 1, 0)
        while k:
            if k & 1:
                part = _ext_add(part, base)
            base = _ext_double(base)
            k >>= 1
        acc = _ext_add(acc, part)
    return _from_ext(acc)


def is_identity(pt) -> bool:
    return pt[0] == 0 and pt[1] == 1


def valid_point(pt) -> bool:
    """curve.py:56-67: on curve, non-identity, and in the prime-order subgroup."""
    if is_identity(pt) or not is_on_curve(pt):
        return False
    cleared = mul_raw(pt, COFACTOR)
    if is_identity(cleared):
        return False
    return mul_raw(cleared, pow(COFACTOR, -1, N)) == pt


# ---- codec (point.py:150-214, te_affine_point.py:297-316, vrf/codec.py:39-45) ------


def point_to_string(pt) -> bytes:
    x, y = pt
    out = bytearray(y.to_bytes(32, "little"))
    if x > (-x) % P:
        out[31] |= 0x80
    return bytes(out)


def string_to_point(data: bytes):
    """Decode without the subgroup check; raises ValueError on a bad encoding."""
    if len(data) != 32:
        raise ValueError("Invalid point encoding")
    sign = data[31] >> 7
    y = int.from_bytes(data[:31] + bytes([data[31] & 0x7F]), "little")
    if y >= P:
        raise ValueError("Invalid point encoding")
    lhs = (1 - y * y) % P
    rhs = (A - D * y * y) % P
    if rhs == 0:
        raise ValueError("Invalid point encoding")
    x = fr_sqrt(lhs * pow(rhs, -1, P))
    if x is None:
        raise ValueError("Invalid point encoding")
    lo, hi = sorted((x, (-x) % P))
    return (hi if sign else lo, y)


def dec_point(data: bytes):
    """vrf/codec.py:39-45: decode + valid nonidentity subgroup point."""
    if len(data) != 32:
        raise ValueError("point must be exactly 32 bytes")
    pt = string_to_point(data)
    if not valid_point(pt):
        raise ValueError("point is not a valid nonidentity subgroup point")
    return pt


# ---- hash to curve (curve.py:110-230, te_curve.py:48-95, te_affine_point.py:195-290) ---


def expand_message_xmd_sha512(msg: bytes, dst: bytes, len_in_bytes: int) -> bytes:
    b_in_bytes, r_in_bytes = 64, 48  # NB: the suite sets expand_len (Z_pad) to 48 (specs/bandersnatch.py:80)
    ell = -(-len_in_bytes // b_in_bytes)
    dst_prime = dst + bytes([len(dst)])
    msg_prime = bytes(r_in_bytes) + msg + len_in_bytes.to_bytes(2, "big") + b"\x00" + dst_prime
    b0 = hashlib.sha512(msg_prime).digest()
    bs = [hashlib.sha512(b0 + b"\x01" + dst_prime).digest()]
    for i in range(2, ell + 1):
        x = bytes(p ^ q for p, q in zip(b0, bs[-1], strict=True))
        bs.append(hashlib.sha512(x + bytes([i]) + dst_prime).digest())
    return b"".join(bs)[:len_in_bytes]


def expand_message_xof_shake128(msg: bytes, dst: bytes, len_in_bytes: int) -> bytes:
    dst_prime = dst + bytes([len(dst)])
    return hashlib.shake_128(msg + len_in_bytes.to_bytes(2, "big") + dst_prime).digest(len_in_bytes)


def hash_to_field(suite: Suite, msg: bytes, count: int) -> list[int]:
    length = 48 * count
    if suite.hash_name == "sha512":
        u = expand_message_xmd_sha512(msg, suite.dst, length)
    else:
        u = expand_message_xof_shake128(msg, suite.dst, length)
    return [int.from_bytes(u[48 * i : 48 * i + 48], "big") % P for i in range(count)]


_MONT_DENOM_INV = pow((A - D) % P, -1, P)
MONT_A = 2 * (A + D) * _MONT_DENOM_INV % P
MONT_B = 4 * _MONT_DENOM_INV % P
ELL2_Z = 5


def map_to_curve_ell2(u: int):
    """Elligator 2 onto the Montgomery model (te_curve.py:48-95)."""
    a_over_b = MONT_A * pow(MONT_B, -1, P) % P
    inv_b2 = pow(MONT_B * MONT_B % P, -1, P)
    tv1 = ELL2_Z * u * u % P
    if tv1 == P - 1:
        tv1 = 0
    x1 = (-a_over_b) * pow(tv1 + 1, -1, P) % P
    gx1 = ((x1 + a_over_b) * x1 + inv_b2) % P * x1 % P
    x2 = (-x1 - a_over_b) % P
    gx2 = tv1 * gx1 % P
    e2 = fr_is_square(gx1)
    x, y2 = (x1, gx1) if e2 else (x2, gx2)
    y = fr_sqrt(y2)
    e3 = (y % 2) == 1
    if e2 ^ e3:
        y = (-y) % P
    return (x * MONT_B % P, y * MONT_B % P)


def from_mont(s: int, t: int):
    """Montgomery (s, t) -> twisted Edwards (te_affine_point.py:262-290)."""
    tv1 = (s + 1) % P
    tv2 = tv1 * t % P
    tv2 = pow(tv2, -1, P) if tv2 else 0
    v = tv2 * tv1 % P * s % P
    w = tv2 * t % P * ((s - 1) % P) % P
    if tv2 == 0:
        w = 1
    return (v, w)


def encode_to_curve(suite: Suite, alpha: bytes, salt: bytes = b""):
    """Elligator2 random-oracle variant (te_affine_point.py:213-222)."""
    u0, u1 = hash_to_field(suite, salt + alpha, 2)
    q0 = from_mont(*map_to_curve_ell2(u0))
    q1 = from_mont(*map_to_curve_ell2(u1))
    return mul_raw(_affine_add_general(q0, q1), COFACTOR)


def _affine_add_general(p1, p2):
    return _from_ext(_ext_add(_to_ext(p1), _to_ext(p2)))
