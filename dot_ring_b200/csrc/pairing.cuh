// BLS12-381 optimal-ate pairing check on Montgomery Fq limbs (host + device).
//
// Stands in for what the reference obtains from blst through dot_ring/ring_proof/pcs/pairing.py:24-31
// (`blst.PT(P2_Affine, P1_Affine)` = Miller loop, `PT.finalverify` = equality after the final
// exponentiation), used by KZG.verify / batch_verify_linear_preconverted (pcs/kzg.py:194-338).
// Only the verdict  e(A1, B1) == e(A2, B2)  is observable in the reference, so the tower, the line
// scaling and the exponent multiple below are ours:
//   Fq2 = Fq[u]/(u^2+1),  Fq6 = Fq2[v]/(v^3 - xi), xi = 1+u,  Fq12 = Fq6[w]/(w^2 - v)   (w^6 = xi)
//   M-type twist E'/Fq2: y^2 = x^3 + 4 xi, untwist (x', y') -> (x'/w^2, y'/w^3).
//   Lines are scaled by elements of Fq4 (killed by the final exponentiation) into the sparse form
//   l = a + b v + c v w  (positions 0, 1, 4 of the six Fq2 coordinates).
//   Final exponentiation: easy part (p^6-1)(p^2+1); hard part via
//   3 (p^4 - p^2 + 1)/r = (x-1)^2 (x+p) (x^2 + p^2 - 1) + 3   (checked numerically in tests/test_host_math.py),
//   five exponentiations by the curve parameter x = -0xd201000000010000.
// One thread evaluates one check; the verifier kernels call pairing_product_is_one() per proof or per batch.
#pragma once
#include "g1.cuh"

namespace dr {

struct Fq2 {
    Fq c0, c1;
    DR_HD static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    DR_HD static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    DR_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    DR_HD bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
};
DR_HD Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
DR_HD Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
DR_HD Fq2 fq2_neg(const Fq2& a) { return {a.c0.neg(), a.c1.neg()}; }
DR_HD Fq2 fq2_conj(const Fq2& a) { return {a.c0, a.c1.neg()}; }
DR_HD Fq2 fq2_dbl(const Fq2& a) { return {a.c0.dbl(), a.c1.dbl()}; }
DR_HD Fq2 fq2_mul_fq(const Fq2& a, const Fq& k) { return {a.c0 * k, a.c1 * k}; }
DR_HD Fq2 fq2_mul_xi(const Fq2& a) { return {a.c0 - a.c1, a.c0 + a.c1}; }  // * (1 + u)
DR_HD_COLD Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
    Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    return {t0 - t1, (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1};
}
DR_HD_COLD Fq2 fq2_sqr(const Fq2& a) {
    Fq t = a.c0 * a.c1;
    return {(a.c0 + a.c1) * (a.c0 - a.c1), t.dbl()};
}
DR_HD_COLD Fq2 fq2_inv(const Fq2& a) {
    Fq d = (a.c0.sqr() + a.c1.sqr()).inv();
    return {a.c0 * d, (a.c1 * d).neg()};
}
DR_HD_COLD Fq2 fq2_pow(const Fq2& a, const uint32_t* e, int nlimbs) {
    Fq2 acc = Fq2::one();
#pragma unroll 1
    for (int i = nlimbs - 1; i >= 0; i--)
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            acc = fq2_sqr(acc);
            if ((e[i] >> b) & 1) acc = fq2_mul(acc, a);
        }
    return acc;
}

struct Fq6 {
    Fq2 c0, c1, c2;
    DR_HD static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    DR_HD static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    DR_HD bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
};
DR_HD Fq6 operator+(const Fq6& a, const Fq6& b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
DR_HD Fq6 operator-(const Fq6& a, const Fq6& b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
DR_HD Fq6 fq6_neg(const Fq6& a) { return {fq2_neg(a.c0), fq2_neg(a.c1), fq2_neg(a.c2)}; }
DR_HD Fq6 fq6_mul_v(const Fq6& a) { return {fq2_mul_xi(a.c2), a.c0, a.c1}; }
DR_HD_COLD Fq6 fq6_mul(const Fq6& a, const Fq6& b) {
    Fq2 t0 = fq2_mul(a.c0, b.c0), t1 = fq2_mul(a.c1, b.c1), t2 = fq2_mul(a.c2, b.c2);
    Fq2 c0 = t0 + fq2_mul_xi(fq2_mul(a.c1 + a.c2, b.c1 + b.c2) - t1 - t2);
    Fq2 c1 = fq2_mul(a.c0 + a.c1, b.c0 + b.c1) - t0 - t1 + fq2_mul_xi(t2);
    Fq2 c2 = fq2_mul(a.c0 + a.c2, b.c0 + b.c2) - t0 - t2 + t1;
    return {c0, c1, c2};
}
DR_HD_COLD Fq6 fq6_inv(const Fq6& a) {
    Fq2 t0 = fq2_sqr(a.c0) - fq2_mul_xi(fq2_mul(a.c1, a.c2));
    Fq2 t1 = fq2_mul_xi(fq2_sqr(a.c2)) - fq2_mul(a.c0, a.c1);
    Fq2 t2 = fq2_sqr(a.c1) - fq2_mul(a.c0, a.c2);
    Fq2 d = fq2_inv(fq2_mul(a.c0, t0) + fq2_mul_xi(fq2_mul(a.c2, t1) + fq2_mul(a.c1, t2)));
    return {fq2_mul(t0, d), fq2_mul(t1, d), fq2_mul(t2, d)};
}

struct Fq12 {
    Fq6 c0, c1;
    DR_HD static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    DR_HD bool is_one() const { return c0 == Fq6::one() && c1 == Fq6::zero(); }
};
DR_HD_COLD Fq12 fq12_mul(const Fq12& a, const Fq12& b) {
    Fq6 t0 = fq6_mul(a.c0, b.c0), t1 = fq6_mul(a.c1, b.c1);
    return {t0 + fq6_mul_v(t1), fq6_mul(a.c0 + a.c1, b.c0 + b.c1) - t0 - t1};
}
DR_HD_COLD Fq12 fq12_sqr(const Fq12& a) {
    Fq6 ab = fq6_mul(a.c0, a.c1);
    Fq6 c0 = fq6_mul(a.c0 + a.c1, a.c0 + fq6_mul_v(a.c1)) - ab - fq6_mul_v(ab);
    return {c0, ab + ab};
}
DR_HD Fq12 fq12_conj(const Fq12& a) { return {a.c0, fq6_neg(a.c1)}; }
DR_HD_COLD Fq12 fq12_inv(const Fq12& a) {
    Fq6 t = fq6_inv(fq6_mul(a.c0, a.c0) - fq6_mul_v(fq6_mul(a.c1, a.c1)));
    return {fq6_mul(a.c0, t), fq6_neg(fq6_mul(a.c1, t))};
}
// f * (o0 + o1 v + o4 v w): the sparse line value
DR_HD_COLD Fq12 fq12_mul_by_014(const Fq12& f, const Fq2& o0, const Fq2& o1, const Fq2& o4) {
    // (f0 + f1 w)(s0 + s1 w), s0 = o0 + o1 v, s1 = o4 v:  c0 = f0 s0 + v f1 s1,  c1 = f0 s1 + f1 s0
    Fq6 s0{o0, o1, Fq2::zero()};
    Fq6 f0s0 = fq6_mul(f.c0, s0), f1s0 = fq6_mul(f.c1, s0);
    // x * (o4 v) = (xi x2 o4, x0 o4, x1 o4)
    Fq6 f0s1{fq2_mul_xi(fq2_mul(f.c0.c2, o4)), fq2_mul(f.c0.c0, o4), fq2_mul(f.c0.c1, o4)};
    Fq6 f1s1{fq2_mul_xi(fq2_mul(f.c1.c2, o4)), fq2_mul(f.c1.c0, o4), fq2_mul(f.c1.c1, o4)};
    return {f0s0 + fq6_mul_v(f1s1), f0s1 + f1s0};
}

// Frobenius constants: gw = xi^((p-1)/6), gv1 = gw^2 = xi^((p-1)/3), gv2 = gw^4 (computed once on the host).
struct PairingConsts {
    Fq2 gw, gv1, gv2;
};
inline PairingConsts pairing_consts_host() {
    // (p - 1) / 6 by schoolbook division of the little-endian limbs
    uint32_t e[12];
    for (int i = 0; i < 12; i++) e[i] = FqTag::mod(i);
    e[0] -= 1;  // p is odd, no borrow
    uint64_t rem = 0;
    for (int i = 11; i >= 0; i--) {
        uint64_t cur = (rem << 32) | e[i];
        e[i] = (uint32_t)(cur / 6);
        rem = cur % 6;
    }
    PairingConsts k;
    Fq2 xi{Fq::one(), Fq::one()};
    k.gw = fq2_pow(xi, e, 12);
    k.gv1 = fq2_sqr(k.gw);
    k.gv2 = fq2_sqr(k.gv1);
    return k;
}
DR_HD_COLD Fq12 fq12_frob(const Fq12& a, const PairingConsts& k) {
    Fq6 c0{fq2_conj(a.c0.c0), fq2_mul(fq2_conj(a.c0.c1), k.gv1), fq2_mul(fq2_conj(a.c0.c2), k.gv2)};
    Fq6 c1{fq2_mul(fq2_conj(a.c1.c0), k.gw), fq2_mul(fq2_mul(fq2_conj(a.c1.c1), k.gv1), k.gw), fq2_mul(fq2_mul(fq2_conj(a.c1.c2), k.gv2), k.gw)};
    return {c0, c1};
}

struct G2Affine {  // on the twist, Montgomery Fq2 coordinates; never infinity here (SRS points)
    Fq2 x, y;
};
// 192-byte zcash uncompressed: x.c1 | x.c0 | y.c1 | y.c0 (srs.py:80-88).  false on a malformed encoding.
DR_HD_COLD bool g2_decode_uncompressed(G2Affine& out, const uint8_t* in192) {
    if (in192[0] & 0xE0) return false;
    Fq t[4];
    for (int i = 0; i < 4; i++) {
        fq_from_be_bytes_raw(t[i], in192 + 48 * i);
        if (!t[i].is_canonical_raw()) return false;
        t[i] = t[i].to_mont();
    }
    out.x = {t[1], t[0]};
    out.y = {t[3], t[2]};
    Fq2 four_xi = fq2_mul_xi(Fq2{Fq::from_u32(4), Fq::zero()});
    return fq2_sqr(out.y) == fq2_mul(fq2_sqr(out.x), out.x) + four_xi;
}

struct G2Proj {
    Fq2 X, Y, Z;
};
// T <- 2T and f <- f * line_{T,T}(P)
DR_HD_COLD void miller_dbl_step(Fq12& f, G2Proj& T, const Fq& px, const Fq& py) {
    Fq2 XX = fq2_sqr(T.X);
    Fq2 W = fq2_dbl(XX) + XX;
    Fq2 S = fq2_mul(T.Y, T.Z);
    Fq2 YS = fq2_mul(T.Y, S);
    Fq2 B = fq2_mul(T.X, YS);
    Fq2 B4 = fq2_dbl(fq2_dbl(B));
    Fq2 H = fq2_sqr(W) - fq2_dbl(B4);
    Fq2 l0 = fq2_mul(W, T.X) - fq2_dbl(YS);
    Fq2 l1 = fq2_neg(fq2_mul_fq(fq2_mul(W, T.Z), px));
    Fq2 l4 = fq2_mul_fq(fq2_dbl(fq2_mul(S, T.Z)), py);
    f = fq12_mul_by_014(f, l0, l1, l4);
    Fq2 SS = fq2_sqr(S);
    Fq2 YS2 = fq2_sqr(YS);
    T.X = fq2_dbl(fq2_mul(H, S));
    T.Y = fq2_mul(W, B4 - H) - fq2_dbl(fq2_dbl(fq2_dbl(YS2)));
    T.Z = fq2_dbl(fq2_dbl(fq2_dbl(fq2_mul(SS, S))));
}
// T <- T + Q and f <- f * line_{T,Q}(P)
DR_HD_COLD void miller_add_step(Fq12& f, G2Proj& T, const G2Affine& Q, const Fq& px, const Fq& py) {
    Fq2 u = fq2_mul(Q.y, T.Z) - T.Y;
    Fq2 v = fq2_mul(Q.x, T.Z) - T.X;
    Fq2 l0 = fq2_mul(u, Q.x) - fq2_mul(v, Q.y);
    Fq2 l1 = fq2_neg(fq2_mul_fq(u, px));
    Fq2 l4 = fq2_mul_fq(v, py);
    f = fq12_mul_by_014(f, l0, l1, l4);
    Fq2 vv = fq2_sqr(v);
    Fq2 vvv = fq2_mul(v, vv);
    Fq2 R = fq2_mul(vv, T.X);
    Fq2 A = fq2_mul(fq2_sqr(u), T.Z) - vvv - fq2_dbl(R);
    T.X = fq2_mul(v, A);
    T.Y = fq2_mul(u, R - A) - fq2_mul(vvv, T.Y);
    T.Z = fq2_mul(vvv, T.Z);
}

constexpr uint64_t BLS_X_ABS = 0xd201000000010000ULL;

// prod_i f_{|x|,Q_i}(P_i) for n <= 2 pairs, squarings shared.  Pairs whose G1 point is infinity contribute 1.
DR_HD_COLD Fq12 miller_loop_product(const G1Affine* P, const G2Affine* Q, int n) {
    Fq12 f = Fq12::one();
    G2Proj T[2];
    bool live[2] = {false, false};
    for (int i = 0; i < n; i++) {
        live[i] = !P[i].is_inf();
        T[i] = {Q[i].x, Q[i].y, Fq2::one()};
    }
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        f = fq12_sqr(f);
#pragma unroll 1
        for (int i = 0; i < n; i++)
            if (live[i]) miller_dbl_step(f, T[i], P[i].x, P[i].y);
        if ((BLS_X_ABS >> b) & 1) {
#pragma unroll 1
            for (int i = 0; i < n; i++)
                if (live[i]) miller_add_step(f, T[i], Q[i], P[i].x, P[i].y);
        }
    }
    return f;
}

// a^x for the (negative) curve parameter; a must lie in the cyclotomic subgroup (inverse = conjugate)
DR_HD_COLD Fq12 fq12_exp_x(const Fq12& a) {
    Fq12 acc = a;
#pragma unroll 1
    for (int b = 62; b >= 0; b--) {
        acc = fq12_sqr(acc);
        if ((BLS_X_ABS >> b) & 1) acc = fq12_mul(acc, a);
    }
    return fq12_conj(acc);
}

DR_HD_COLD Fq12 final_exponentiation(const Fq12& f, const PairingConsts& k) {
    Fq12 f1 = fq12_mul(fq12_conj(f), fq12_inv(f));          // ^(p^6 - 1)
    Fq12 f2 = fq12_mul(fq12_frob(fq12_frob(f1, k), k), f1);  // ^(p^2 + 1)
    Fq12 a = fq12_mul(fq12_exp_x(f2), fq12_conj(f2));        // ^(x - 1)
    a = fq12_mul(fq12_exp_x(a), fq12_conj(a));               // ^(x - 1)^2
    Fq12 b = fq12_mul(fq12_exp_x(a), fq12_frob(a, k));       // ^(x + p)
    Fq12 c = fq12_mul(fq12_mul(fq12_exp_x(fq12_exp_x(b)), fq12_frob(fq12_frob(b, k), k)), fq12_conj(b));  // ^(x^2 + p^2 - 1)
    return fq12_mul(c, fq12_mul(fq12_sqr(f2), f2));          // * f2^3
}

// e(a1, b1) == e(a2, b2)   <=>   e(a1, b1) * e(-a2, b2) == 1
DR_HD_COLD bool pairing_equal(const G1Affine& a1, const G2Affine& b1, const G1Affine& a2, const G2Affine& b2, const PairingConsts& k) {
    G1Affine P[2] = {a1, a2};
    if (!P[1].is_inf()) P[1].y = P[1].y.neg();
    G2Affine Q[2] = {b1, b2};
    return final_exponentiation(miller_loop_product(P, Q, 2), k).is_one();
}

}  // namespace dr
