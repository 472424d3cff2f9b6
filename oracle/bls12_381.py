"""BLS12-381 restatement (oracle; test infrastructure only).

Stands in for the third-party ``blst`` / ``py_ecc`` calls the reference makes at
  dot_ring/ring_proof/pcs/kzg.py:121-175,194-338   (MSM, codecs, verify)
  dot_ring/ring_proof/pcs/srs.py:42-148            (SRS points)
  dot_ring/ring_proof/pcs/pairing.py:24-31         (Miller loop / finalverify)
  dot_ring/ring_proof/pcs/utils.py:38-58           (compressed conversions)
Encodings follow the zcash BLS12-381 serialisation the reference documents in
tests/utils/rust_serde.py:102-155 (flag bits 0x80 compressed, 0x40 infinity,
0x20 lexicographically-larger y).

Points: G1 Jacobian ``(X, Y, Z)`` over ints, ``None`` is the point at infinity.
G2 affine ``((x0, x1), (y0, y1))`` with Fq2 = Fq[i]/(i^2+1), element c0 + c1*i.
"""

from __future__ import annotations

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
B_G1 = 4
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
ATE_LOOP_COUNT = 0xD201000000010000  # |x| of the BLS parameter (x is negative)

# ----------------------------------------------------------------------------
# G1 (Jacobian)
# ----------------------------------------------------------------------------


def g1_is_on_curve(x: int, y: int) -> bool:
    return (y * y - x * x * x - B_G1) % P == 0


def g1_from_affine(x: int, y: int):
    return (x % P, y % P, 1)


def g1_to_affine(pt):
    """Jacobian -> affine (x, y); None stays None."""
    if pt is None:
        return None
    X, Y, Z = pt
    if Z == 1:
        return (X, Y)
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 % P * zi % P)


def g1_double(pt):
    if pt is None:
        return None
    X, Y, Z = pt
    if Y == 0:
        return None
    A = X * X % P
    Bq = Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def g1_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 % P * Z2Z2 % P
    S2 = Y2 * Z1 % P * Z1Z1 % P
    H = (U2 - U1) % P
    Rr = (S2 - S1) % P
    if H == 0:
        if Rr == 0:
            return g1_double(p1)
        return None
    HH = H * H % P
    HHH = H * HH % P
    V = U1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    Y3 = (Rr * (V - X3) - S1 * HHH) % P
    Z3 = Z1 * Z2 % P * H % P
    return (X3, Y3, Z3)


def g1_add_affine(p1, a):
    """Jacobian + affine (x, y) mixed addition."""
    if a is None:
        return p1
    if p1 is None:
        return (a[0], a[1], 1)
    X1, Y1, Z1 = p1
    x2, y2 = a
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 % P * Z1Z1 % P
    H = (U2 - X1) % P
    Rr = (S2 - Y1) % P
    if H == 0:
        if Rr == 0:
            return g1_double(p1)
        return None
    HH = H * H % P
    HHH = H * HH % P
    V = X1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    Y3 = (Rr * (V - X3) - Y1 * HHH) % P
    Z3 = Z1 * H % P
    return (X3, Y3, Z3)


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P, pt[2])


def g1_mul(pt, k: int):
    """Scalar multiple; k is taken mod r (blst P1.mult semantics on group scalars)."""
    k %= R
    acc = None
    base = pt
    while k:
        if k & 1:
            acc = g1_add(acc, base)
        base = g1_double(base)
        k >>= 1
    return acc


def g1_eq(p1, p2) -> bool:
    return g1_to_affine(p1) == g1_to_affine(p2)


_C_MSM = None


def _load_c_msm():
    """ctypes handle on oracle/c/liboracle_msm.so (plain-C Pippenger), or False when it is not built."""
    global _C_MSM
    if _C_MSM is None:
        import ctypes
        from pathlib import Path

        so = Path(__file__).resolve().parent / "c" / "liboracle_msm.so"
        try:
            lib = ctypes.CDLL(str(so))
            lib.oracle_g1_msm.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
            _C_MSM = lib
        except OSError:
            _C_MSM = False
    return _C_MSM


def g1_msm(points_affine, scalars, window: int | None = None):
    """MSM entry point of the oracle: the C restatement when built (same algorithm, 64-bit limbs), else Python."""
    lib = _load_c_msm()
    n = min(len(points_affine), len(scalars))
    if lib and window is None and n > 0 and all(p is not None for p in points_affine[:n]):
        import ctypes

        pts = b"".join(x.to_bytes(48, "big") + y.to_bytes(48, "big") for x, y in points_affine[:n])
        ks = b"".join((int(k) % R).to_bytes(32, "little") for k in scalars[:n])
        out = ctypes.create_string_buffer(96)
        lib.oracle_g1_msm(pts, ks, n, out)
        return g1_decompress(out.raw)
    return g1_msm_python(points_affine, scalars, window)


def g1_msm_python(points_affine, scalars, window: int | None = None):
    """Pippenger bucket MSM over affine (x, y) bases (None = infinity).

    Restates the semantics of blst ``P1_Affines.mult_pippenger`` as the
    reference calls it (kzg.py:149,170-173,295-296,332-333): scalars are Python
    ints that may exceed r and are group scalars (taken mod r); result is a
    projective point.  The window schedule is free (group result is unique).
    """
    n = min(len(points_affine), len(scalars))
    if n == 0:
        return None
    ks = [int(s) % R for s in scalars[:n]]
    if window is None:
        window = 3 if n < 32 else (n.bit_length() * 69 // 100 + 2)
        window = max(2, min(window, 14))
    c = window
    nwin = (255 + c - 1) // c
    mask = (1 << c) - 1
    total = None
    for w in range(nwin - 1, -1, -1):
        if total is not None:
            for _ in range(c):
                total = g1_double(total)
        buckets = [None] * (mask + 1)
        shift = w * c
        for i in range(n):
            d = (ks[i] >> shift) & mask
            if d:
                buckets[d] = g1_add_affine(buckets[d], points_affine[i])
        run = None
        acc = None
        for d in range(mask, 0, -1):
            if buckets[d] is not None:
                run = g1_add(run, buckets[d])
            if run is not None:
                acc = g1_add(acc, run)
        total = g1_add(total, acc)
    return total


# ---- zcash codecs (kzg.py:121-144, srs.py:57-66) ---------------------------


def fq_sqrt(a: int):
    """sqrt in Fq (p = 3 mod 4); None if a is a non-residue."""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a else None


def g1_serialize(pt) -> bytes:
    """96-byte uncompressed zcash encoding (blst ``P1.serialize``)."""
    a = g1_to_affine(pt)
    if a is None:
        return bytes([0x40]) + bytes(95)
    return a[0].to_bytes(48, "big") + a[1].to_bytes(48, "big")


def g1_compress(pt) -> bytes:
    """48-byte compressed zcash encoding (blst ``P1.compress``)."""
    a = g1_to_affine(pt)
    if a is None:
        return bytes([0xC0]) + bytes(47)
    x, y = a
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > (P - 1) // 2:
        out[0] |= 0x20
    return bytes(out)


def g1_decompress(data: bytes, check_subgroup: bool = False):
    """Inverse of :func:`g1_compress`; raises ValueError on a bad encoding.

    Mirrors ``blst.P1_Affine(bytes)`` raising on malformed input, which the
    reference maps to ``ValueError("invalid BLS12-381 G1 encoding")``
    (kzg.py:138-144).  Returns a Jacobian point or None for infinity.
    """
    if len(data) == 96 and not (data[0] & 0x80):
        if data[0] & 0x40:
            if any(data[1:]) or data[0] != 0x40:
                raise ValueError("invalid BLS12-381 G1 encoding")
            return None
        if data[0] & 0x20:
            raise ValueError("invalid BLS12-381 G1 encoding")
        x = int.from_bytes(data[:48], "big")
        y = int.from_bytes(data[48:], "big")
        if x >= P or y >= P or not g1_is_on_curve(x, y):
            raise ValueError("invalid BLS12-381 G1 encoding")
        pt = (x, y, 1)
        if check_subgroup and g1_mul_raw(pt, R) is not None:
            raise ValueError("invalid BLS12-381 G1 encoding")
        return pt
    if len(data) != 48:
        raise ValueError("invalid BLS12-381 G1 encoding")
    flags = data[0]
    if not flags & 0x80:
        raise ValueError("invalid BLS12-381 G1 encoding")
    if flags & 0x40:
        if flags != 0xC0 or any(data[1:]):
            raise ValueError("invalid BLS12-381 G1 encoding")
        return None
    x = int.from_bytes(bytes([flags & 0x1F]) + data[1:], "big")
    if x >= P:
        raise ValueError("invalid BLS12-381 G1 encoding")
    y = fq_sqrt(x * x * x + B_G1)
    if y is None:
        raise ValueError("invalid BLS12-381 G1 encoding")
    if (y > (P - 1) // 2) != bool(flags & 0x20):
        y = P - y
    pt = (x, y, 1)
    if check_subgroup and g1_mul_raw(pt, R) is not None:
        raise ValueError("invalid BLS12-381 G1 encoding")
    return pt


def g1_mul_raw(pt, k: int):
    """Scalar multiple without reducing k mod r (used for the subgroup check)."""
    acc = None
    base = pt
    while k:
        if k & 1:
            acc = g1_add(acc, base)
        base = g1_double(base)
        k >>= 1
    return acc


# ----------------------------------------------------------------------------
# Fq2 and G2 (affine)
# ----------------------------------------------------------------------------


def fq2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fq2_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return ((a0 * b0 - a1 * b1) % P, (a0 * b1 + a1 * b0) % P)


def fq2_sqr(a):
    a0, a1 = a
    return ((a0 + a1) * (a0 - a1) % P, 2 * a0 * a1 % P)


def fq2_inv(a):
    a0, a1 = a
    d = pow((a0 * a0 + a1 * a1) % P, -1, P)
    return (a0 * d % P, (-a1 * d) % P)


def fq2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def fq2_scalar(a, k: int):
    return (a[0] * k % P, a[1] * k % P)


FQ2_ZERO = (0, 0)
FQ2_ONE = (1, 0)
B_G2 = (4, 4)


def fq2_pow(a, e: int):
    out = FQ2_ONE
    base = a
    while e:
        if e & 1:
            out = fq2_mul(out, base)
        base = fq2_sqr(base)
        e >>= 1
    return out


def fq2_sqrt(a):
    """Square root in Fq2 (p = 3 mod 4), Algorithm 9 of eprint 2012/685; None if none."""
    if a == FQ2_ZERO:
        return FQ2_ZERO
    a1 = fq2_pow(a, (P - 3) // 4)
    alpha = fq2_mul(fq2_sqr(a1), a)
    a0 = fq2_mul(fq2_pow(alpha, P), alpha)
    if a0 == (P - 1, 0):
        return None
    x0 = fq2_mul(a1, a)
    if alpha == (P - 1, 0):
        x = fq2_mul((0, 1), x0)
    else:
        b = fq2_pow(fq2_add(FQ2_ONE, alpha), (P - 1) // 2)
        x = fq2_mul(b, x0)
    return x if fq2_sqr(x) == (a[0] % P, a[1] % P) else None


def g2_is_on_curve(pt) -> bool:
    x, y = pt
    return fq2_sub(fq2_sqr(y), fq2_add(fq2_mul(fq2_sqr(x), x), B_G2)) == FQ2_ZERO


def g2_double(pt):
    if pt is None:
        return None
    x, y = pt
    if y == FQ2_ZERO:
        return None
    lam = fq2_mul(fq2_scalar(fq2_sqr(x), 3), fq2_inv(fq2_scalar(y, 2)))
    x3 = fq2_sub(fq2_sqr(lam), fq2_scalar(x, 2))
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x, x3)), y)
    return (x3, y3)


def g2_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 == y2:
            return g2_double(p1)
        return None
    lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_sqr(lam), x1), x2)
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_neg(pt):
    if pt is None:
        return None
    return (pt[0], fq2_neg(pt[1]))


def g2_mul(pt, k: int):
    k %= R
    acc = None
    base = pt
    while k:
        if k & 1:
            acc = g2_add(acc, base)
        base = g2_double(base)
        k >>= 1
    return acc


def g2_serialize(pt) -> bytes:
    """192-byte uncompressed zcash encoding: x.c1 | x.c0 | y.c1 | y.c0 (srs.py:80-88)."""
    if pt is None:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = pt
    return x1.to_bytes(48, "big") + x0.to_bytes(48, "big") + y1.to_bytes(48, "big") + y0.to_bytes(48, "big")


def g2_from_uncompressed(data: bytes):
    if len(data) != 192:
        raise ValueError("invalid BLS12-381 G2 encoding")
    x1 = int.from_bytes(data[0:48], "big")
    x0 = int.from_bytes(data[48:96], "big")
    y1 = int.from_bytes(data[96:144], "big")
    y0 = int.from_bytes(data[144:192], "big")
    pt = ((x0, x1), (y0, y1))
    if max(x0, x1, y0, y1) >= P or not g2_is_on_curve(pt):
        raise ValueError("invalid BLS12-381 G2 encoding")
    return pt


def g2_decompress(data: bytes):
    """96-byte compressed zcash G2 (c1 | c0, flags in byte 0)."""
    if len(data) != 96 or not data[0] & 0x80:
        raise ValueError("invalid BLS12-381 G2 encoding")
    flags = data[0]
    if flags & 0x40:
        return None
    x1 = int.from_bytes(bytes([flags & 0x1F]) + data[1:48], "big")
    x0 = int.from_bytes(data[48:96], "big")
    if x0 >= P or x1 >= P:
        raise ValueError("invalid BLS12-381 G2 encoding")
    x = (x0, x1)
    y = fq2_sqrt(fq2_add(fq2_mul(fq2_sqr(x), x), B_G2))
    if y is None:
        raise ValueError("invalid BLS12-381 G2 encoding")
    y0, y1 = y
    larger = (y1 > (P - 1) // 2) if y1 != 0 else (y0 > (P - 1) // 2)
    if larger != bool(flags & 0x20):
        y = fq2_neg(y)
    return (x, y)


# ----------------------------------------------------------------------------
# Fq12 = Fq[w]/(w^12 - 2 w^6 + 2), ate pairing (py_ecc-style formulation)
# ----------------------------------------------------------------------------

FQ12_ONE = (1,) + (0,) * 11
FQ12_ZERO = (0,) * 12


def fq12_mul(a, b):
    t = [0] * 23
    for i in range(12):
        ai = a[i]
        if ai:
            for j in range(12):
                t[i + j] += ai * b[j]
    # w^12 = 2 w^6 - 2
    for k in range(22, 11, -1):
        v = t[k]
        if v:
            t[k - 6] += 2 * v
            t[k - 12] -= 2 * v
    return tuple(x % P for x in t[:12])


def fq12_sqr(a):
    return fq12_mul(a, a)


def _poly_deg(p):
    d = len(p) - 1
    while d >= 0 and p[d] == 0:
        d -= 1
    return d


def fq12_inv(a):
    """Inverse by the extended Euclidean algorithm on polynomials over Fq."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low = list(a) + [0]
    high = [2, 0, 0, 0, 0, 0, (-2) % P, 0, 0, 0, 0, 0, 1]
    while _poly_deg(low) > 0:
        # r = high // low (rounded poly division)
        dl = _poly_deg(low)
        temp = list(high)
        rq = [0] * 13
        inv_lead = pow(low[dl], -1, P)
        for i in range(_poly_deg(temp) - dl, -1, -1):
            q = temp[dl + i] * inv_lead % P
            rq[i] = q
            if q:
                for c in range(dl + 1):
                    temp[c + i] = (temp[c + i] - q * low[c]) % P
        nm = list(hm)
        new = list(high)
        for i in range(13):
            if lm[i] or low[i]:
                for j in range(13 - i):
                    if rq[j]:
                        nm[i + j] = (nm[i + j] - lm[i] * rq[j]) % P
                        new[i + j] = (new[i + j] - low[i] * rq[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, P)
    return tuple(x * inv0 % P for x in lm[:12])


def fq12_pow(a, e: int):
    out = FQ12_ONE
    base = a
    while e:
        if e & 1:
            out = fq12_mul(out, base)
        base = fq12_sqr(base)
        e >>= 1
    return out


def _fq2_embed(a, shift: int):
    """Fq2 element c0 + c1*i (i = w^6 - 1) times w^shift, as an Fq12 tuple (shift < 6)."""
    out = [0] * 12
    out[shift] = (a[0] - a[1]) % P
    out[shift + 6] = a[1] % P
    return tuple(out)


_W = (0, 1) + (0,) * 10
_W_INV = fq12_inv(_W)
_W_INV3 = fq12_mul(fq12_mul(_W_INV, _W_INV), _W_INV)


def _line(lam, x1, y1, px: int, py: int):
    """Line through the twisted G2 point with Fq2 slope lam, evaluated at G1 (px, py).

    With the twist (x, y) -> (x / w^2, y / w^3) the line value is
    -py + (lam*px) * w^-1 + (y1 - lam*x1) * w^-3.
    """
    t1 = fq12_mul(_fq2_embed(fq2_scalar(lam, px), 0), _W_INV)
    t3 = fq12_mul(_fq2_embed(fq2_sub(y1, fq2_mul(lam, x1)), 0), _W_INV3)
    out = [(t1[i] + t3[i]) % P for i in range(12)]
    out[0] = (out[0] - py) % P
    return tuple(out)


def miller_loop(q_affine, p_affine):
    """Ate Miller loop f_{|x|,Q}(P); stands in for ``blst.PT(P2_Affine, P1_Affine)``
    (pairing.py:24-26).  Any fixed non-degenerate bilinear map yields the same
    verdicts in ``final_verify``; the GT value itself is never serialised."""
    if q_affine is None or p_affine is None:
        return FQ12_ONE
    px, py = p_affine
    qx, qy = q_affine
    rx, ry = qx, qy
    f = FQ12_ONE
    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        lam = fq2_mul(fq2_scalar(fq2_sqr(rx), 3), fq2_inv(fq2_scalar(ry, 2)))
        f = fq12_mul(fq12_sqr(f), _line(lam, rx, ry, px, py))
        nx = fq2_sub(fq2_sqr(lam), fq2_scalar(rx, 2))
        ry = fq2_sub(fq2_mul(lam, fq2_sub(rx, nx)), ry)
        rx = nx
        if (ATE_LOOP_COUNT >> i) & 1:
            lam = fq2_mul(fq2_sub(qy, ry), fq2_inv(fq2_sub(qx, rx)))
            f = fq12_mul(f, _line(lam, rx, ry, px, py))
            nx = fq2_sub(fq2_sub(fq2_sqr(lam), rx), qx)
            ry = fq2_sub(fq2_mul(lam, fq2_sub(rx, nx)), ry)
            rx = nx
    return f


_FINAL_EXP = (P**12 - 1) // R


def final_exponentiation(f):
    return fq12_pow(f, _FINAL_EXP)


def final_verify(lhs, rhs) -> bool:
    """``blst.PT.finalverify`` (pairing.py:29-31): final_exp(lhs) == final_exp(rhs)."""
    return final_exponentiation(fq12_mul(lhs, fq12_inv(rhs))) == FQ12_ONE


def pairing_check_eq(p1_affine, q1_affine, p2_affine, q2_affine) -> bool:
    """e(p1, q1) == e(p2, q2)."""
    return final_verify(miller_loop(q1_affine, p1_affine), miller_loop(q2_affine, p2_affine))
