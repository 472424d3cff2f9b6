// Warp-cooperative pairing check: one 32-thread block evaluates  e(P1, Q1) * e(P2, Q2) == 1  for FIXED G2 points Q1, Q2.
//
// The one-thread pairing of pairing.cuh spends ~20 k dependent 381-bit multiplications (50 ms on a B200); the verifier's G2
// points are always the two SRS points of the verifier key (pcs/kzg.py:260-263), so
//   * the Miller-loop lines are precomputed per verifier key (68 per G2 point: 63 doublings + 5 additions), and a step is only
//     f <- f^2 * l_1(P1) * l_2(P2);
//   * an Fq12 product is split over the lanes: 18 Fq2 products (Karatsuba over the tower) on 18 lanes, 9 lanes recombine
//     the three Fq6 products, 6 lanes write the result -- about 4 multiplication latencies instead of 54.
// G1 points are taken in XYZZ form: the line  l0 + (l1' x) v + (l4' y) v w  is scaled by ZZ*ZZZ (an Fq factor, killed by the
// final exponentiation), so no inversion is needed before the loop.
// Code is written in the phase style of rt.cuh (DR_THREAD_LOOP / DR_BLOCK_SYNC), all shared state in shared memory.
#pragma once
#include "pairing.cuh"
#include "rt.cuh"

namespace dr {

constexpr int MILLER_LINES = 68;  // per G2 point: for bit 62..0 of |x|: doubling line, plus an addition line where the bit is set

struct LineCoeffs {
    Fq2 l0, l1, l4;  // l1, l4 still to be multiplied by x(P), y(P)
};

// Host-side precomputation (verifier-key set-up).  Same step formulas as pairing.cuh, without the evaluation point.
inline void miller_precompute_lines(const G2Affine& Q, LineCoeffs* out) {
    G2Proj T{Q.x, Q.y, Fq2::one()};
    int n = 0;
    for (int b = 62; b >= 0; b--) {
        {  // doubling
            Fq2 XX = fq2_sqr(T.X);
            Fq2 W = fq2_dbl(XX) + XX;
            Fq2 S = fq2_mul(T.Y, T.Z);
            Fq2 YS = fq2_mul(T.Y, S);
            Fq2 B = fq2_mul(T.X, YS);
            Fq2 B4 = fq2_dbl(fq2_dbl(B));
            Fq2 H = fq2_sqr(W) - fq2_dbl(B4);
            out[n++] = {fq2_mul(W, T.X) - fq2_dbl(YS), fq2_neg(fq2_mul(W, T.Z)), fq2_dbl(fq2_mul(S, T.Z))};
            Fq2 SS = fq2_sqr(S);
            Fq2 YS2 = fq2_sqr(YS);
            T.X = fq2_dbl(fq2_mul(H, S));
            T.Y = fq2_mul(W, B4 - H) - fq2_dbl(fq2_dbl(fq2_dbl(YS2)));
            T.Z = fq2_dbl(fq2_dbl(fq2_dbl(fq2_mul(SS, S))));
        }
        if ((BLS_X_ABS >> b) & 1) {  // addition of Q
            Fq2 u = fq2_mul(Q.y, T.Z) - T.Y;
            Fq2 v = fq2_mul(Q.x, T.Z) - T.X;
            out[n++] = {fq2_mul(u, Q.x) - fq2_mul(v, Q.y), fq2_neg(u), v};
            Fq2 vv = fq2_sqr(v);
            Fq2 vvv = fq2_mul(v, vv);
            Fq2 R = fq2_mul(vv, T.X);
            Fq2 A = fq2_mul(fq2_sqr(u), T.Z) - vvv - fq2_dbl(R);
            T.X = fq2_mul(v, A);
            T.Y = fq2_mul(u, R - A) - fq2_mul(vvv, T.Y);
            T.Z = fq2_mul(vvv, T.Z);
        }
    }
}

// shared-memory working set of one pairing block
struct PairingWarpState {
    Fq12 f, t0, t1, t2, t3;  // accumulators / temporaries (6 Fq2 each in tower order c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2)
    Fq12 line;
    Fq2 prod[18];
    Fq2 z[9];
    Fq coord[2][3];  // per pair: ZZ*ZZZ, X*ZZZ, Y*ZZ
    uint32_t live[2];
    uint32_t verdict;
};

DR_HD Fq2* fq12_coeffs(Fq12* a) { return &a->c0.c0; }
DR_HD const Fq2* fq12_coeffs(const Fq12* a) { return &a->c0.c0; }

// dst = a * b (dst may alias a or b)
// out of line: called ~600 times per pairing; inlining it everywhere costs registers and instruction-cache misses
DR_HD_COLD void fq12w_mul(const BlockCtx& ctx, PairingWarpState* st, Fq12* dst, const Fq12* a, const Fq12* b) {
    const Fq2* ca = fq12_coeffs(a);
    const Fq2* cb = fq12_coeffs(b);
    DR_THREAD_LOOP(t, ctx) {
        if (t < 18) {
            uint32_t s = t / 6, q = t % 6;
            // operand selection: Fq12 level (s) then Fq6 level (q)
            uint32_t i0 = q < 3 ? q : (q == 3 ? 1 : 0), i1 = q < 3 ? 3 : (q == 4 ? 1 : 2);  // q<3: single index; else pair (i0, i1)
            Fq2 x, y;
            if (s == 0) {
                x = ca[i0];
                y = cb[i0];
            } else if (s == 1) {
                x = ca[3 + i0];
                y = cb[3 + i0];
            } else {
                x = ca[i0] + ca[3 + i0];
                y = cb[i0] + cb[3 + i0];
            }
            if (q >= 3) {
                if (s == 0) {
                    x = x + ca[i1];
                    y = y + cb[i1];
                } else if (s == 1) {
                    x = x + ca[3 + i1];
                    y = y + cb[3 + i1];
                } else {
                    x = x + ca[i1] + ca[3 + i1];
                    y = y + cb[i1] + cb[3 + i1];
                }
            }
            st->prod[t] = fq2_mul(x, y);
        }
    }
    DR_BLOCK_SYNC();
    DR_THREAD_LOOP(t, ctx) {
        if (t < 9) {
            uint32_t s = t / 3, j = t % 3;
            const Fq2* p = st->prod + 6 * s;
            Fq2 r;
            if (j == 0) r = p[0] + fq2_mul_xi(p[3] - p[1] - p[2]);
            else if (j == 1) r = p[4] - p[0] - p[1] + fq2_mul_xi(p[2]);
            else r = p[5] - p[0] - p[2] + p[1];
            st->z[t] = r;
        }
    }
    DR_BLOCK_SYNC();
    Fq2* cd = fq12_coeffs(dst);
    DR_THREAD_LOOP(t, ctx) {
        if (t < 6) {
            const Fq2* z = st->z;  // z[0..2] = A0*B0, z[3..5] = A1*B1, z[6..8] = (A0+A1)(B0+B1)
            Fq2 r;
            if (t == 0) r = z[0] + fq2_mul_xi(z[5]);
            else if (t == 1) r = z[1] + z[3];
            else if (t == 2) r = z[2] + z[4];
            else r = z[6 + (t - 3)] - z[t - 3] - z[3 + (t - 3)];
            cd[t] = r;
        }
    }
    DR_BLOCK_SYNC();
}

DR_HD void fq12w_copy(const BlockCtx& ctx, Fq12* dst, const Fq12* src) {
    DR_THREAD_LOOP(t, ctx) {
        if (t < 6) fq12_coeffs(dst)[t] = fq12_coeffs(src)[t];
    }
    DR_BLOCK_SYNC();
}
DR_HD void fq12w_conj(const BlockCtx& ctx, Fq12* dst, const Fq12* src) {
    DR_THREAD_LOOP(t, ctx) {
        if (t < 6) fq12_coeffs(dst)[t] = t < 3 ? fq12_coeffs(src)[t] : fq2_neg(fq12_coeffs(src)[t]);
    }
    DR_BLOCK_SYNC();
}
// Frobenius: coefficient t is conjugated and scaled by 1, gv1, gv2, gw, gw*gv1, gw*gv2
DR_HD_COLD void fq12w_frob(const BlockCtx& ctx, Fq12* dst, const Fq12* src, const PairingConsts& k) {
    DR_THREAD_LOOP(t, ctx) {
        if (t < 6) {
            Fq2 c = fq2_conj(fq12_coeffs(src)[t]);
            if (t == 1 || t == 4) c = fq2_mul(c, k.gv1);
            if (t == 2 || t == 5) c = fq2_mul(c, k.gv2);
            if (t >= 3) c = fq2_mul(c, k.gw);
            fq12_coeffs(dst)[t] = c;
        }
    }
    DR_BLOCK_SYNC();
}
// dst = a^x (x = -|x|; a in the cyclotomic subgroup).  tmp must differ from dst and a.
DR_HD_COLD void fq12w_exp_x(const BlockCtx& ctx, PairingWarpState* st, Fq12* dst, const Fq12* a, Fq12* tmp) {
    fq12w_copy(ctx, tmp, a);
    for (int b = 62; b >= 0; b--) {
        fq12w_mul(ctx, st, tmp, tmp, tmp);
        if ((BLS_X_ABS >> b) & 1) fq12w_mul(ctx, st, tmp, tmp, a);
    }
    fq12w_conj(ctx, dst, tmp);
}

// The whole check for one block.  P[i] in XYZZ (infinity allowed: that pair contributes 1); P[1] is negated by the caller
// when the check is an equality  e(P0, Q0) == e(P1', Q1).  lines: [2][MILLER_LINES] in device memory.  Result in st->verdict.
DR_HD_COLD void pairing_product_is_one_warp(const BlockCtx& ctx, PairingWarpState* st, const G1* P, const LineCoeffs* lines, const PairingConsts& k) {
    DR_THREAD_LOOP(t, ctx) {
        if (t < 2) {
            const G1& p = P[t];
            st->live[t] = p.is_inf() ? 0u : 1u;
            st->coord[t][0] = p.ZZ * p.ZZZ;
            st->coord[t][1] = p.X * p.ZZZ;
            st->coord[t][2] = p.Y * p.ZZ;
        }
        if (t < 6) fq12_coeffs(&st->f)[t] = t == 0 ? Fq2::one() : Fq2::zero();
    }
    DR_BLOCK_SYNC();
    int li = 0;
    for (int b = 62; b >= 0; b--) {
        fq12w_mul(ctx, st, &st->f, &st->f, &st->f);
        const int nsteps = ((BLS_X_ABS >> b) & 1) ? 2 : 1;
        for (int sidx = 0; sidx < nsteps; sidx++, li++) {
            for (int pr = 0; pr < 2; pr++) {
                if (!st->live[pr]) continue;
                const LineCoeffs& L = lines[pr * MILLER_LINES + li];
                DR_THREAD_LOOP(t, ctx) {
                    if (t < 6) {
                        Fq2 c = Fq2::zero();
                        if (t == 0) c = fq2_mul_fq(L.l0, st->coord[pr][0]);
                        if (t == 1) c = fq2_mul_fq(L.l1, st->coord[pr][1]);
                        if (t == 4) c = fq2_mul_fq(L.l4, st->coord[pr][2]);
                        fq12_coeffs(&st->line)[t] = c;
                    }
                }
                DR_BLOCK_SYNC();
                fq12w_mul(ctx, st, &st->f, &st->f, &st->line);
            }
        }
    }
    // ---- final exponentiation (pairing.cuh: final_exponentiation) ----
    // f1 = conj(f) * f^-1 ; the inverse is serial (one lane)
    DR_THREAD_LOOP(t, ctx) {
        if (t == 0) st->t0 = fq12_inv(st->f);
    }
    DR_BLOCK_SYNC();
    fq12w_conj(ctx, &st->t1, &st->f);
    fq12w_mul(ctx, st, &st->t1, &st->t1, &st->t0);  // t1 = f1
    fq12w_frob(ctx, &st->t0, &st->t1, k);
    fq12w_frob(ctx, &st->t0, &st->t0, k);
    fq12w_mul(ctx, st, &st->f, &st->t0, &st->t1);  // f = f2 (cyclotomic)
    // a = f2^(x-1)
    fq12w_exp_x(ctx, st, &st->t0, &st->f, &st->t3);
    fq12w_conj(ctx, &st->t1, &st->f);
    fq12w_mul(ctx, st, &st->t0, &st->t0, &st->t1);  // t0 = a
    // a = a^(x-1)
    fq12w_exp_x(ctx, st, &st->t1, &st->t0, &st->t3);
    fq12w_conj(ctx, &st->t2, &st->t0);
    fq12w_mul(ctx, st, &st->t0, &st->t1, &st->t2);  // t0 = a = f2^((x-1)^2)
    // b = a^(x+p)
    fq12w_exp_x(ctx, st, &st->t1, &st->t0, &st->t3);
    fq12w_frob(ctx, &st->t2, &st->t0, k);
    fq12w_mul(ctx, st, &st->t0, &st->t1, &st->t2);  // t0 = b
    // c = b^(x^2) * b^(p^2) * b^-1
    fq12w_exp_x(ctx, st, &st->t1, &st->t0, &st->t3);
    fq12w_exp_x(ctx, st, &st->t2, &st->t1, &st->t3);  // t2 = b^(x^2)
    fq12w_frob(ctx, &st->t1, &st->t0, k);
    fq12w_frob(ctx, &st->t1, &st->t1, k);
    fq12w_mul(ctx, st, &st->t2, &st->t2, &st->t1);
    fq12w_conj(ctx, &st->t1, &st->t0);
    fq12w_mul(ctx, st, &st->t2, &st->t2, &st->t1);  // t2 = c
    // result = c * f2^3
    fq12w_mul(ctx, st, &st->t1, &st->f, &st->f);
    fq12w_mul(ctx, st, &st->t1, &st->t1, &st->f);
    fq12w_mul(ctx, st, &st->t2, &st->t2, &st->t1);
    DR_THREAD_LOOP(t, ctx) {
        if (t == 0) st->verdict = st->t2.is_one() ? 1u : 0u;
    }
    DR_BLOCK_SYNC();
}

}  // namespace dr
