// BLS12-381 G1 (y^2 = x^3 + 4) group law on Montgomery Fq limbs.
//
// Stands in for blst's P1 / P1_Affine arithmetic the reference reaches through
// dot_ring/ring_proof/pcs/kzg.py:147-175 (mult_pippenger), :121-144 (codecs) and
// dot_ring/ring_proof/pcs/utils.py:38-46.  Accumulators use extended Jacobian ("XYZZ")
// coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, infinity <=> ZZ == 0.  Mixed addition is
// 8M+2S, full addition 12M+2S, doubling 6M+4S (EFD madd-2008-s / add-2008-s / dbl-2008-s-1).
// All exceptional cases (infinity operands, P == Q, P == -Q) are handled so results are exact for
// every input, not only generic ones.
#pragma once
#include "fp.cuh"

namespace dr {

struct G1Affine {  // 96 bytes; infinity encoded as (0, 0) (not a curve point since b = 4)
    Fq x, y;
    DR_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    DR_HD static G1Affine inf() { return {Fq::zero(), Fq::zero()}; }
};

struct G1 {  // XYZZ
    Fq X, Y, ZZ, ZZZ;
    DR_HD bool is_inf() const { return ZZ.is_zero(); }
    DR_HD static G1 inf() { return {Fq::zero(), Fq::zero(), Fq::zero(), Fq::zero()}; }
    DR_HD static G1 from_affine(const G1Affine& a) {
        if (a.is_inf()) return inf();
        return {a.x, a.y, Fq::one(), Fq::one()};
    }
};

DR_HD_COLD G1 g1_dbl_affine(const G1Affine& a) {
    if (a.is_inf() || a.y.is_zero()) return G1::inf();
    Fq U = a.y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = a.x * V;
    Fq xx = a.x.sqr();
    Fq M = xx.dbl() + xx;
    G1 r;
    r.X = M.sqr() - S.dbl();
    r.Y = M * (S - r.X) - W * a.y;
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

DR_HD_COLD G1 g1_dbl(const G1& p) {
    if (p.is_inf() || p.Y.is_zero()) return G1::inf();
    Fq U = p.Y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = p.X * V;
    Fq xx = p.X.sqr();
    Fq M = xx.dbl() + xx;
    G1 r;
    r.X = M.sqr() - S.dbl();
    r.Y = M * (S - r.X) - W * p.Y;
    r.ZZ = V * p.ZZ;
    r.ZZZ = W * p.ZZZ;
    return r;
}

// acc += a (mixed).  `neg` adds -a instead.
DR_HD void g1_madd(G1& acc, const G1Affine& a, bool neg = false) {
    if (a.is_inf()) return;
    Fq ay = neg ? a.y.neg() : a.y;
    if (acc.is_inf()) {
        acc.X = a.x;
        acc.Y = ay;
        acc.ZZ = Fq::one();
        acc.ZZZ = Fq::one();
        return;
    }
    Fq P = a.x * acc.ZZ - acc.X;
    Fq R = ay * acc.ZZZ - acc.Y;
    if (P.is_zero()) {
        if (R.is_zero()) {
            G1Affine t{a.x, ay};
            acc = g1_dbl_affine(t);
        } else {
            acc = G1::inf();
        }
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = acc.X * PP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    acc.Y = R * (Q - X3) - acc.Y * PPP;
    acc.X = X3;
    acc.ZZ = acc.ZZ * PP;
    acc.ZZZ = acc.ZZZ * PPP;
}

DR_HD_COLD void g1_add(G1& acc, const G1& b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) {
        acc = b;
        return;
    }
    Fq U1 = acc.X * b.ZZ;
    Fq S1 = acc.Y * b.ZZZ;
    Fq P = b.X * acc.ZZ - U1;
    Fq R = b.Y * acc.ZZZ - S1;
    if (P.is_zero()) {
        if (R.is_zero()) {
            acc = g1_dbl(acc);
        } else {
            acc = G1::inf();
        }
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = U1 * PP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    acc.Y = R * (Q - X3) - S1 * PPP;
    acc.X = X3;
    acc.ZZ = acc.ZZ * b.ZZ * PP;
    acc.ZZZ = acc.ZZZ * b.ZZZ * PPP;
}

// k * P for canonical little-endian limbs k (8 limbs), fixed 4-bit windows
DR_HD_COLD G1 g1_mul_limbs(const G1& p, const uint32_t* k) {
    if (p.is_inf()) return G1::inf();
    G1 tab[16];
    tab[0] = G1::inf();
    tab[1] = p;
#pragma unroll 1
    for (int i = 2; i < 16; i++) {
        if (i & 1) {
            tab[i] = tab[i - 1];
            g1_add(tab[i], p);
        } else {
            tab[i] = g1_dbl(tab[i >> 1]);
        }
    }
    G1 acc = G1::inf();
#pragma unroll 1
    for (int i = 7; i >= 0; i--) {
#pragma unroll 1
        for (int sft = 28; sft >= 0; sft -= 4) {
            if (!acc.is_inf()) acc = g1_dbl(g1_dbl(g1_dbl(g1_dbl(acc))));
            uint32_t d = (k[i] >> sft) & 15;
            if (d) g1_add(acc, tab[d]);
        }
    }
    return acc;
}

DR_HD G1 g1_neg(const G1& p) {
    G1 r = p;
    r.Y = p.Y.neg();
    return r;
}

DR_HD_COLD G1Affine g1_to_affine(const G1& p) {
    if (p.is_inf()) return G1Affine::inf();
    Fq t = (p.ZZ * p.ZZZ).inv();
    Fq zz_inv = t * p.ZZZ;
    Fq zzz_inv = t * p.ZZ;
    return {p.X * zz_inv, p.Y * zzz_inv};
}

DR_HD bool g1_affine_on_curve(const G1Affine& a) {
    Fq four = Fq::from_u32(4);
    return a.y.sqr() == a.x.sqr() * a.x + four;
}

// ---- zcash codecs (kzg.py:121-144; blst P1.serialize / P1.compress / P1_Affine(bytes)) ---------
// y is "lexicographically larger" iff y > (p-1)/2  <=>  2y > p  (canonical integers).
DR_HD bool fq_is_lex_larger(const Fq& y_mont) {
    Fq y = y_mont.from_mont();
    // compare 2y with p: y > (p-1)/2
    uint32_t carry = 0;
    uint32_t d[Fq::N];
    for (int i = 0; i < Fq::N; i++) {
        d[i] = (y.v[i] << 1) | carry;
        carry = y.v[i] >> 31;
    }
    if (carry) return true;
    return Fq::geq_mod(d);  // 2y >= p, and 2y != p since p is odd
}

DR_HD void g1_serialize(uint8_t* out96, const G1Affine& a) {
    if (a.is_inf()) {
        for (int i = 0; i < 96; i++) out96[i] = 0;
        out96[0] = 0x40;
        return;
    }
    fq_to_be_bytes_raw(out96, a.x.from_mont());
    fq_to_be_bytes_raw(out96 + 48, a.y.from_mont());
}

DR_HD void g1_compress(uint8_t* out48, const G1Affine& a) {
    if (a.is_inf()) {
        for (int i = 0; i < 48; i++) out48[i] = 0;
        out48[0] = 0xC0;
        return;
    }
    fq_to_be_bytes_raw(out48, a.x.from_mont());
    out48[0] |= 0x80;
    if (fq_is_lex_larger(a.y)) out48[0] |= 0x20;
}

// sqrt in Fq: p = 3 mod 4, candidate a^((p+1)/4)
DR_HD_COLD bool fq_sqrt(Fq& out, const Fq& a) {
    // (p+1)/4 little-endian limbs
    constexpr uint32_t e[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};
    uint32_t ee[12];
    for (int i = 0; i < 12; i++) ee[i] = e[i];
    Fq s = a.pow(ee, 12);
    out = s;
    return s.sqr() == a;
}

// Decode 48-byte compressed or 96-byte uncompressed zcash G1.  Returns false on a malformed
// encoding (blst raises; the reference turns that into ValueError("invalid BLS12-381 G1 encoding")).
// No subgroup check, matching blst's P1_Affine(bytes) constructor.
DR_HD_COLD bool g1_decode(G1Affine& out, const uint8_t* in, int len) {
    if (len != 48 && len != 96) return false;
    uint8_t flags = in[0];
    bool compressed = (flags & 0x80) != 0;
    if (compressed != (len == 48)) return false;
    if (flags & 0x40) {
        if ((flags & 0x3F) != 0) return false;
        for (int i = 1; i < len; i++)
            if (in[i]) return false;
        out = G1Affine::inf();
        return true;
    }
    uint8_t buf[48];
    for (int i = 0; i < 48; i++) buf[i] = in[i];
    buf[0] &= 0x1F;
    Fq xr;
    fq_from_be_bytes_raw(xr, buf);
    if (!xr.is_canonical_raw()) return false;
    Fq x = xr.to_mont();
    if (!compressed) {
        if (flags & 0x20) return false;
        Fq yr;
        fq_from_be_bytes_raw(yr, in + 48);
        if (!yr.is_canonical_raw()) return false;
        out = {x, yr.to_mont()};
        return g1_affine_on_curve(out);
    }
    Fq y;
    if (!fq_sqrt(y, x.sqr() * x + Fq::from_u32(4))) return false;
    if (fq_is_lex_larger(y) != ((flags & 0x20) != 0)) y = y.neg();
    out = {x, y};
    return true;
}

}  // namespace dr
