// Bandersnatch (twisted Edwards, a = -5) over Fr: group law, codec, subgroup check, Elligator 2.
//
// Replaces, for batched use on the device:
//   dot_ring/curve/native_field/bandersnatch_te.pyx:127-174  (extended add / double)
//   dot_ring/curve/native_field/bandersnatch_te.pyx:421-477  (Tonelli-Shanks in Fr)
//   dot_ring/curve/native_field/bandersnatch_te.pyx:480-816  (small simultaneous multiplications)
//   dot_ring/curve/twisted_edwards/te_affine_point.py:69-316 (affine law, from_mont, x-recover)
//   dot_ring/curve/twisted_edwards/te_curve.py:48-95         (Elligator 2 map)
//   dot_ring/curve/point.py:150-214, dot_ring/vrf/codec.py:39-45, dot_ring/curve/curve.py:56-67
//   dot_ring/curve/glv.py:128-189, specs/bandersnatch.py:65-67,177-286 (GLV: endomorphism, scalar split)
// Points are unique group elements, so the multiplication schedule is ours: full-length multiplications of subgroup points go
// through the endomorphism like the reference's (two 128-bit halves, half the doublings) but with 4-bit windows and shared
// doublings (Straus) instead of its 2-bit joint windows; fixed bases use window tables without doublings.
#pragma once
#include "fp.cuh"

namespace dr {

struct TEAffine {
    Fr x, y;  // Montgomery
    DR_HD static TEAffine identity() { return {Fr::zero(), Fr::one()}; }
    DR_HD bool is_identity() const { return x.is_zero() && y == Fr::one(); }
    DR_HD bool operator==(const TEAffine& o) const { return x == o.x && y == o.y; }
};

struct TEExt {
    Fr X, Y, Z, T;
    DR_HD static TEExt identity() { return {Fr::zero(), Fr::one(), Fr::one(), Fr::zero()}; }
    DR_HD static TEExt from_affine(const TEAffine& a) { return {a.x, a.y, Fr::one(), a.x * a.y}; }
    DR_HD bool is_identity() const { return X.is_zero() && Y == Z; }
};

DR_HD Fr te_d() {
    constexpr uint32_t c[8] = DR_TE_D;
    Fr r;
    for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}
DR_HD Fr fr_mul5(const Fr& a) {
    Fr a2 = a.dbl();
    return a2.dbl() + a;
}

// add-2008-hwcd with a = -5:  H = B - a*A = B + 5A
DR_HD TEExt te_add(const TEExt& p, const TEExt& q) {
    Fr A = p.X * q.X;
    Fr B = p.Y * q.Y;
    Fr C = p.T * te_d() * q.T;
    Fr D = p.Z * q.Z;
    Fr E = (p.X + p.Y) * (q.X + q.Y) - A - B;
    Fr F = D - C;
    Fr G = D + C;
    Fr H = B + fr_mul5(A);
    return {E * F, G * H, F * G, E * H};
}

// mixed: q affine with precomputed td = d * x * y
struct TEPre {
    Fr x, y, td;
    DR_HD static TEPre from_affine(const TEAffine& a) { return {a.x, a.y, a.x * a.y * te_d()}; }
};
DR_HD TEExt te_madd(const TEExt& p, const TEPre& q) {
    Fr A = p.X * q.x;
    Fr B = p.Y * q.y;
    Fr C = p.T * q.td;
    Fr E = (p.X + p.Y) * (q.x + q.y) - A - B;
    Fr F = p.Z - C;
    Fr G = p.Z + C;
    Fr H = B + fr_mul5(A);
    return {E * F, G * H, F * G, E * H};
}

// dbl-2008-hwcd: D = a*A = -5A
DR_HD TEExt te_dbl(const TEExt& p) {
    Fr A = p.X.sqr();
    Fr B = p.Y.sqr();
    Fr C = p.Z.sqr().dbl();
    Fr D = fr_mul5(A).neg();
    Fr E = (p.X + p.Y).sqr() - A - B;
    Fr G = D + B;
    Fr F = G - C;
    Fr H = D - B;
    return {E * F, G * H, F * G, E * H};
}

DR_HD TEExt te_neg(const TEExt& p) { return {p.X.neg(), p.Y, p.Z, p.T.neg()}; }
DR_HD TEAffine te_neg(const TEAffine& p) { return {p.x.neg(), p.y}; }

DR_HD TEAffine te_to_affine(const TEExt& p) {
    Fr zi = p.Z.inv();
    return {p.X * zi, p.Y * zi};
}

DR_HD bool te_ext_eq_affine(const TEExt& p, const TEAffine& a) { return p.X == a.x * p.Z && p.Y == a.y * p.Z; }

DR_HD bool te_on_curve(const TEAffine& a) {
    Fr x2 = a.x.sqr(), y2 = a.y.sqr();
    return y2 - fr_mul5(x2) == Fr::one() + te_d() * x2 * y2;
}

// k * P for a raw little-endian scalar of `nlimbs` limbs (no reduction), fixed 4-bit windows.
DR_HD_COLD TEExt te_mul_raw(const TEAffine& p, const uint32_t* k, int nlimbs) {
    TEExt tab[16];
    tab[0] = TEExt::identity();
    tab[1] = TEExt::from_affine(p);
#pragma unroll 1
    for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? te_add(tab[i - 1], tab[1]) : te_dbl(tab[i >> 1]);
    TEExt acc = TEExt::identity();
    bool started = false;
#pragma unroll 1
    for (int i = nlimbs - 1; i >= 0; i--) {
#pragma unroll 1
        for (int s = 28; s >= 0; s -= 4) {
            if (started) {
                acc = te_dbl(te_dbl(te_dbl(te_dbl(acc))));
            }
            uint32_t d = (k[i] >> s) & 15;
            if (d) {
                acc = started ? te_add(acc, tab[d]) : tab[d];
                started = true;
            }
        }
    }
    return acc;
}

// k * P for a point with a precomputed window table (8-bit windows over 256 bits: tab[256 * w + d] = d * 2^(8w) * P in
// affine form, d >= 1; built once per base point by TeFixedTableBody): at most 32 mixed additions, no doublings.
constexpr int TE_FIXED_WINDOWS = 32;
constexpr int TE_FIXED_ENTRIES = TE_FIXED_WINDOWS * 256;
DR_HD_COLD TEExt te_mul_fixed(const TEPre* tab, const uint32_t* k) {
    TEExt acc = TEExt::identity();
#pragma unroll 1
    for (int w = 0; w < TE_FIXED_WINDOWS; w++) {
        uint32_t d = (k[w >> 2] >> (8 * (w & 3))) & 255;
        if (d) acc = te_madd(acc, tab[256 * w + d]);
    }
    return acc;
}
// two conversions around one field inversion
DR_HD void te_to_affine2(const TEExt& p, const TEExt& q, TEAffine& pa, TEAffine& qa) {
    Fr zi = (p.Z * q.Z).inv();
    Fr pzi = zi * q.Z, qzi = zi * p.Z;
    pa = {p.X * pzi, p.Y * pzi};
    qa = {q.X * qzi, q.Y * qzi};
}

// sum_i k_i * P_i with shared doublings (Straus), n <= 3, raw 8-limb scalars (< subgroup order).
DR_HD_COLD TEExt te_msm_small(const TEAffine* pts, const uint32_t (*ks)[8], int n) {
    TEExt tab[3][16];
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        tab[j][0] = TEExt::identity();
        tab[j][1] = TEExt::from_affine(pts[j]);
#pragma unroll 1
        for (int i = 2; i < 16; i++) tab[j][i] = (i & 1) ? te_add(tab[j][i - 1], tab[j][1]) : te_dbl(tab[j][i >> 1]);
    }
    TEExt acc = TEExt::identity();
    bool started = false;
#pragma unroll 1
    for (int i = 7; i >= 0; i--) {
#pragma unroll 1
        for (int s = 28; s >= 0; s -= 4) {
            if (started) acc = te_dbl(te_dbl(te_dbl(te_dbl(acc))));
#pragma unroll 1
            for (int j = 0; j < n; j++) {
                uint32_t d = (ks[j][i] >> s) & 15;
                if (d) {
                    acc = started ? te_add(acc, tab[j][d]) : tab[j][d];
                    started = true;
                }
            }
        }
    }
    return acc;
}

// [n]P == O the long way (kept as the cross-check of te_in_prime_subgroup in the unit tests)
DR_HD_COLD bool te_in_prime_subgroup_by_order(const TEAffine& p) {
    constexpr uint32_t n[8] = DR_FN_RAW;
    uint32_t k[8];
    for (int i = 0; i < 8; i++) k[i] = n[i];
    const TEExt r = te_mul_raw(p, k, 8);
    return !r.Z.is_zero() && r.is_identity();
}
// Membership in the prime-order subgroup for a point ON the curve, by 2-descent instead of a 253-bit multiplication.
// a and d are both non-squares, so a d is a square, the curve has full rational 2-torsion and E(Fr) = Z2 x Z2 x Z_n: the prime
// subgroup is exactly 2E, and P is in 2E iff the three descent values of its Montgomery image u = (1 + y) / (1 - y),
//   B u,  B (u - u+),  B (u - u-)      (u+-: the roots of u^2 + A u + 1; their product is a square already)
// are squares.  In Edwards coordinates, with s = sqrt(a d):
//   (a - d)(a - d y^2)  and  2 (1 - y) ((a - s) - (d - s) y)   are both non-zero squares.
// Two Legendre symbols (binary Jacobi algorithm) replace ~2700 field multiplications; the reference multiplies by the order
// (curve.py:56-67), the verdict is the same.  x = 0 holds the identity (in) and the 2-torsion point (0, -1) (out).
// ---- GLV (curve/glv.py:128-189; specs/bandersnatch.py:65-67,177-286) ---------------------------------------------------------
// Bandersnatch has an endomorphism psi with psi(P) = lambda P on the prime subgroup, so k P = k1 P + k2 psi(P) with |k1|, |k2| < 2^127
// and a joint multiplication does half the doublings.  The split uses the short lattice basis v1 = (a1, b1), v2 = (a2, -a1) of
// {(x, y): x + y lambda = 0 mod n}:  c1 = floor(k a1 / n), c2 = floor(k b1 / n) (Barrett, exact to within one),
//   k1 = k - c1 a1 - c2 a2,   k2 = c2 a1 - c1 b1.
// Any (k1, k2) with k1 + k2 lambda = k is as good as the reference's (the group element is the same); magnitudes and signs out.
DR_HD void limbs_mul(uint32_t* out, const uint32_t* a, int na, const uint32_t* b, int nb) {  // out[na + nb] = a * b
    for (int i = 0; i < na + nb; i++) out[i] = 0;
    for (int i = 0; i < na; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < nb; j++) {
            uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        out[i + nb] = (uint32_t)carry;
    }
}
// k: 8 raw limbs below the group order.  k1, k2: 4-limb magnitudes (upper limbs of the 8-limb outputs are zero), neg1 / neg2: signs.
DR_HD_COLD void te_glv_split(const uint32_t* k, uint32_t* k1, bool& neg1, uint32_t* k2, bool& neg2) {
    constexpr uint32_t A1[4] = DR_TE_GLV_A1, B1[4] = DR_TE_GLV_B1, A2[4] = DR_TE_GLV_A2, G1[5] = DR_TE_GLV_G1, G2[4] = DR_TE_GLV_G2;
    uint32_t a1[4], b1[4], a2[4], g1[5], g2[4];
    for (int i = 0; i < 4; i++) {
        a1[i] = A1[i];
        b1[i] = B1[i];
        a2[i] = A2[i];
        g2[i] = G2[i];
    }
    for (int i = 0; i < 5; i++) g1[i] = G1[i];
    uint32_t wide[13], c1[5], c2[4];
    limbs_mul(wide, k, 8, g1, 5);
    for (int i = 0; i < 5; i++) c1[i] = wide[8 + i];  // < 2^128: c1[4] == 0
    limbs_mul(wide, k, 8, g2, 4);
    for (int i = 0; i < 4; i++) c2[i] = wide[8 + i];
    uint32_t p[9], q[9], acc[9];
    // k1 = k - c1 a1 - c2 a2  (9-limb two's complement)
    limbs_mul(p, c1, 4, a1, 4);
    p[8] = 0;
    limbs_mul(q, c2, 4, a2, 4);
    q[8] = 0;
    {
        uint64_t borrow = 0;
        for (int i = 0; i < 9; i++) {
            uint64_t t = (uint64_t)(i < 8 ? k[i] : 0u) - p[i] - borrow;
            acc[i] = (uint32_t)t;
            borrow = (t >> 32) & 1;
        }
        borrow = 0;
        for (int i = 0; i < 9; i++) {
            uint64_t t = (uint64_t)acc[i] - q[i] - borrow;
            acc[i] = (uint32_t)t;
            borrow = (t >> 32) & 1;
        }
    }
    auto magnitude = [](const uint32_t* v, uint32_t* out, bool& neg) {
        neg = (v[8] >> 31) != 0;
        uint64_t carry = 1;
        for (int i = 0; i < 8; i++) {
            uint32_t w = v[i];
            if (neg) {
                uint64_t t = (uint64_t)(~w) + carry;
                w = (uint32_t)t;
                carry = t >> 32;
            }
            out[i] = w;
        }
    };
    magnitude(acc, k1, neg1);
    // k2 = c2 a1 - c1 b1
    limbs_mul(p, c2, 4, a1, 4);
    p[8] = 0;
    limbs_mul(q, c1, 4, b1, 4);
    q[8] = 0;
    {
        uint64_t borrow = 0;
        for (int i = 0; i < 9; i++) {
            uint64_t t = (uint64_t)p[i] - q[i] - borrow;
            acc[i] = (uint32_t)t;
            borrow = (t >> 32) & 1;
        }
    }
    magnitude(acc, k2, neg2);
}
// psi(P) in extended coordinates for an affine P of the prime subgroup (glv.py:165-189: x' = c (1 - y^2)(y^2 - b), y' = b (y^2 + b) x y,
// z' = (y^2 - b) x y); the identity maps to the identity.
DR_HD TEExt te_endomorphism(const TEAffine& p) {
    constexpr uint32_t b_c[8] = DR_TE_GLV_B, c_c[8] = DR_TE_GLV_C;
    Fr b, c;
    for (int i = 0; i < 8; i++) {
        b.v[i] = b_c[i];
        c.v[i] = c_c[i];
    }
    if (p.x.is_zero()) return TEExt::identity();
    const Fr y2 = p.y.sqr(), xy = p.x * p.y;
    const Fr h = y2 - b;
    const Fr xp = c * (Fr::one() - y2) * h, yp = b * (y2 + b) * xy, zp = h * xy;
    return {xp * zp, yp * zp, zp.sqr(), xp * yp};
}

// sum_j k_j P_j for up to three points in EXTENDED coordinates and scalars of `nlimbs` limbs: Straus with shared doublings, 4-bit windows
template <int MAXN = 3>
DR_HD_COLD TEExt te_straus_ext(const TEExt* pts, const uint32_t (*ks)[8], int n, int nlimbs) {
    TEExt tab[MAXN][16];
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        tab[j][0] = TEExt::identity();
        tab[j][1] = pts[j];
#pragma unroll 1
        for (int i = 2; i < 16; i++) tab[j][i] = (i & 1) ? te_add(tab[j][i - 1], tab[j][1]) : te_dbl(tab[j][i >> 1]);
    }
    TEExt acc = TEExt::identity();
    bool started = false;
#pragma unroll 1
    for (int i = nlimbs - 1; i >= 0; i--) {
#pragma unroll 1
        for (int s = 28; s >= 0; s -= 4) {
            if (started) acc = te_dbl(te_dbl(te_dbl(te_dbl(acc))));
#pragma unroll 1
            for (int j = 0; j < n; j++) {
                uint32_t d = (ks[j][i] >> s) & 15;
                if (d) {
                    acc = started ? te_add(acc, tab[j][d]) : tab[j][d];
                    started = true;
                }
            }
        }
    }
    return acc;
}
// k P for a point of the prime subgroup and a scalar below the group order, through the endomorphism: 127 doublings instead of 252
DR_HD_COLD TEExt te_mul_glv(const TEAffine& p, const uint32_t* k) {
    uint32_t ks[2][8];
    bool n1, n2;
    te_glv_split(k, ks[0], n1, ks[1], n2);
    TEExt pts[2] = {TEExt::from_affine(n1 ? te_neg(p) : p), te_endomorphism(p)};
    if (n2) pts[1] = te_neg(pts[1]);
    return te_straus_ext<2>(pts, ks, 2, 4);
}

// k P + c Q: P affine in the prime subgroup with a full-length scalar k (split by the endomorphism), Q any point with a scalar of at
// most 128 bits -- the shape of every verification equation (s * input - c * output): one Straus pass over 32 windows
DR_HD_COLD TEExt te_glv_straus2(const TEAffine& p, const uint32_t* k, const TEExt& q, const uint32_t* c128) {
    uint32_t ks[3][8];
    bool n1, n2;
    te_glv_split(k, ks[0], n1, ks[1], n2);
    for (int i = 0; i < 8; i++) ks[2][i] = i < 4 ? c128[i] : 0u;
    TEExt pts[3] = {TEExt::from_affine(n1 ? te_neg(p) : p), te_endomorphism(p), q};
    if (n2) pts[1] = te_neg(pts[1]);
    return te_straus_ext<3>(pts, ks, 3, 4);
}
// a P + b Q + c R with P, Q affine in the prime subgroup (full-length scalars a, b, split by the endomorphism) and a 128-bit scalar c
// for R: five tables, ONE pass over 32 windows (the Tiny / Thin verification equation with the delinearisation folded into the scalars)
DR_HD_COLD TEExt te_glv_straus3(const TEAffine& p, const uint32_t* a, const TEAffine& q, const uint32_t* b, const TEExt& r, const uint32_t* c128) {
    uint32_t ks[5][8];
    bool n[4];
    te_glv_split(a, ks[0], n[0], ks[1], n[1]);
    te_glv_split(b, ks[2], n[2], ks[3], n[3]);
    for (int i = 0; i < 8; i++) ks[4][i] = i < 4 ? c128[i] : 0u;
    TEExt pts[5] = {TEExt::from_affine(n[0] ? te_neg(p) : p), te_endomorphism(p), TEExt::from_affine(n[2] ? te_neg(q) : q), te_endomorphism(q), r};
    if (n[1]) pts[1] = te_neg(pts[1]);
    if (n[3]) pts[3] = te_neg(pts[3]);
    return te_straus_ext<5>(pts, ks, 5, 4);
}

DR_HD bool te_in_prime_subgroup(const TEAffine& p) {
    constexpr uint32_t amd_c[8] = DR_TE_A_MINUS_D, ams_c[8] = DR_TE_A_MINUS_S, dms_c[8] = DR_TE_D_MINUS_S;
    if (p.x.is_zero()) return p.y == Fr::one();
    Fr amd, ams, dms;
    for (int i = 0; i < 8; i++) {
        amd.v[i] = amd_c[i];
        ams.v[i] = ams_c[i];
        dms.v[i] = dms_c[i];
    }
    const Fr one = Fr::one();
    const Fr c1 = amd * (fr_mul5(one).neg() - te_d() * p.y.sqr());
    if (c1.legendre() != 1) return false;
    const Fr c2 = (one - p.y).dbl() * (ams - dms * p.y);
    return c2.legendre() == 1;
}

// ---- square roots in Fr (2-adicity 32) ---------------------------------------------------------
// Returns false when `a` is a non-residue.  Which of the two roots comes back is unspecified; every
// caller normalises the sign (x-recover orders the candidates, Elligator fixes the parity).
// Split in two so that Elligator 2 can try a and 5a with ONE exponentiation: begin() raises to the odd part of the group
// order, finish() is the Tonelli-Shanks descent over the 2^32 roots of unity.
struct FrSqrtState {
    Fr x, t;  // a^((q+1)/2), a^q   (p - 1 = q * 2^32)
};
DR_HD_COLD FrSqrtState fr_sqrt_begin(const Fr& a) {
    constexpr uint32_t e_c[8] = DR_FR_TS_QM1_HALF;
    uint32_t e[8];
#pragma unroll 1
    for (int i = 0; i < 8; i++) e[i] = e_c[i];
    Fr w = a.pow(e, 8);  // a^((q-1)/2)
    Fr x = a * w;        // a^((q+1)/2)
    return {x, x * w};   // a^q
}
// false: the operand (non-zero) is a non-residue
DR_HD_COLD bool fr_sqrt_finish(Fr& out, const FrSqrtState& st) {
    constexpr uint32_t c_c[8] = DR_FR_TS_C;
    Fr c;
#pragma unroll 1
    for (int i = 0; i < 8; i++) c.v[i] = c_c[i];
    Fr x = st.x, t = st.t;
    int m = 32;
    Fr one = Fr::one();
#pragma unroll 1
    while (t != one) {
        int i = 0;
        Fr t2 = t;
#pragma unroll 1
        while (t2 != one) {
            t2 = t2.sqr();
            i++;
            if (i == m) return false;  // order of t does not divide 2^(m-1): non-residue
        }
        Fr b = c;
#pragma unroll 1
        for (int j = 0; j < m - i - 1; j++) b = b.sqr();
        x = x * b;
        c = b.sqr();
        t = t * c;
        m = i;
    }
    out = x;
    return true;
}
// the state of 5a from the state of a: (5a)^((q+1)/2) = 5^((q+1)/2) a^((q+1)/2), (5a)^q = 5^q a^q
DR_HD FrSqrtState fr_sqrt_times5(const FrSqrtState& st) {
    constexpr uint32_t zh_c[8] = DR_FR_TS_Z_HALF;
    constexpr uint32_t zq_c[8] = DR_FR_TS_C;
    Fr zh, zq;
    for (int i = 0; i < 8; i++) {
        zh.v[i] = zh_c[i];
        zq.v[i] = zq_c[i];
    }
    return {st.x * zh, st.t * zq};
}
DR_HD_COLD bool fr_sqrt(Fr& out, const Fr& a) {
    if (a.is_zero()) {
        out = a;
        return true;
    }
    return fr_sqrt_finish(out, fr_sqrt_begin(a));
}

DR_HD_COLD bool fr_is_square(const Fr& a) {
    if (a.is_zero()) return true;
    constexpr uint32_t e_c[8] = DR_FR_PM1_HALF;
    uint32_t e[8];
    for (int i = 0; i < 8; i++) e[i] = e_c[i];
    return a.pow(e, 8) == Fr::one();
}

// canonical-integer comparison helpers on raw (non-Montgomery) limbs
DR_HD bool raw_gt(const Fr& a, const Fr& b) {
    for (int i = 7; i >= 0; i--)
        if (a.v[i] != b.v[i]) return a.v[i] > b.v[i];
    return false;
}

// ---- 32-byte point codec (point.py:150-214) ----------------------------------------------------
DR_HD void te_encode(uint8_t* out32, const TEAffine& p) {
    Fr xr = p.x.from_mont();
    Fr nxr = p.x.neg().from_mont();
    fr_to_le_bytes_raw(out32, p.y.from_mont());
    if (raw_gt(xr, nxr)) out32[31] |= 0x80;
}

// Decode without the subgroup check.  false <=> ValueError("Invalid point encoding").
DR_HD_COLD bool te_decode(TEAffine& out, const uint8_t* in32) {
    uint8_t buf[32];
    for (int i = 0; i < 32; i++) buf[i] = in32[i];
    bool sign = (buf[31] >> 7) != 0;
    buf[31] &= 0x7F;
    Fr yr;
    fr_from_le_bytes_raw(yr, buf);
    if (!yr.is_canonical_raw()) return false;
    Fr y = yr.to_mont();
    Fr y2 = y.sqr();
    Fr lhs = Fr::one() - y2;                   // 1 - y^2
    Fr rhs = fr_mul5(Fr::one()).neg() - te_d() * y2;  // a - d y^2
    if (rhs.is_zero()) return false;
    Fr x;
    if (!fr_sqrt(x, lhs * rhs.inv())) return false;
    Fr nx = x.neg();
    bool x_is_larger = raw_gt(x.from_mont(), nx.from_mont());
    // candidates ordered (smaller, larger); sign bit selects the larger one
    if (x_is_larger != sign) x = nx;
    out = {x, y};
    return true;
}

// vrf/codec.py:39-45 `dec_point`: decode + nonidentity prime-subgroup check.
DR_HD bool te_decode_checked(TEAffine& out, const uint8_t* in32) {
    if (!te_decode(out, in32)) return false;
    if (out.is_identity()) return false;
    return te_in_prime_subgroup(out);
}

// ---- Elligator 2 (te_curve.py:48-95, te_affine_point.py:262-290) --------------------------------
DR_HD Fr fr_const(const uint32_t (&c)[8]) {
    Fr r;
    for (int i = 0; i < 8; i++) r.v[i] = c[i];
    return r;
}

// u (Montgomery) -> twisted Edwards affine point
DR_HD_COLD TEAffine te_map_to_curve_ell2(const Fr& u) {
    constexpr uint32_t aob_c[8] = DR_ELL2_A_OVER_B;
    constexpr uint32_t ib2_c[8] = DR_ELL2_INV_B2;
    constexpr uint32_t b_c[8] = DR_ELL2_B;
    Fr a_over_b = fr_const(aob_c), inv_b2 = fr_const(ib2_c), mb = fr_const(b_c);
    Fr one = Fr::one();
    Fr tv1 = fr_mul5(u.sqr());  // Z = 5
    const bool exceptional = tv1 == one.neg();
    if (exceptional) tv1 = Fr::zero();
    Fr x1 = a_over_b.neg() * (tv1 + one).inv();
    Fr gx1 = ((x1 + a_over_b) * x1 + inv_b2) * x1;
    Fr x2 = x1.neg() - a_over_b;
    Fr y = Fr::zero();
    bool e2 = true;
    Fr x = x1;
    if (!gx1.is_zero()) {
        // exactly one of gx1, gx2 = 5 u^2 gx1 is a square, and sqrt(gx2) = u sqrt(5 gx1) comes out of the same exponentiation
        FrSqrtState st = fr_sqrt_begin(gx1);
        e2 = fr_sqrt_finish(y, st);
        if (!e2) {
            x = x2;
            if (exceptional) {
                fr_sqrt(y, tv1 * gx1);  // gx2 = 0
            } else {
                Fr r;
                fr_sqrt_finish(r, fr_sqrt_times5(st));
                y = u * r;
            }
        }
    }
    bool e3 = (y.from_mont().v[0] & 1) != 0;
    if (e2 != e3) y = y.neg();
    Fr s = x * mb, t = y * mb;
    // Montgomery (s, t) -> Edwards (v, w)
    Fr tv1b = s + one;
    Fr tv2 = tv1b * t;
    bool degenerate = tv2.is_zero();
    Fr tv2i = tv2.inv();  // 0 -> 0
    Fr v = tv2i * tv1b * s;
    Fr w = tv2i * t * (s - one);
    if (degenerate) w = one;
    return {v, w};
}

// hash_to_field output bytes (48-byte big-endian each) are reduced mod p on the host or device side:
DR_HD Fr fr_from_be48_mod(const uint8_t* in48) {
    uint8_t le[48];
    for (int i = 0; i < 48; i++) le[i] = in48[47 - i];
    return fp_from_le_bytes_mod<Fr>(le, 48);
}

// Elligator2 random-oracle encode: (u0, u1) -> 4 * (map(u0) + map(u1))   (te_affine_point.py:213-222)
DR_HD TEAffine te_encode_to_curve_from_u(const Fr& u0, const Fr& u1) {
    TEAffine q0 = te_map_to_curve_ell2(u0);
    TEAffine q1 = te_map_to_curve_ell2(u1);
    TEExt s = te_add(TEExt::from_affine(q0), TEExt::from_affine(q1));
    s = te_dbl(te_dbl(s));
    return te_to_affine(s);
}

}  // namespace dr
