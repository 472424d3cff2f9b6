// Cooperative Bandersnatch scalar multiplication: eight threads per item.
//
// The provers run one item per thread where the batch is small (a 512-proof shard of an 8-GPU run, a single `prove` call), and
// there a 253-bit variable-base multiplication is ~2700 field multiplications in ONE dependent chain: every multiplication is a
// carry chain of ~300 integer instructions whose latency nothing hides.  The extended-coordinate formulas are two layers of
// independent products (dbl-2008-hwcd: 4 + 4, add-2008-hwcd: 5 + 4), so the lanes of an item compute one product each per layer
// and exchange them through shared memory: a doubling costs two multiplication latencies instead of eight, an addition two
// instead of nine.  Fixed-base multiplications (8-bit windows over a precomputed table, te.cuh te_mul_fixed) are split by
// windows instead: four table additions per lane, then one lane folds the eight partial sums.
//
// Written in the phase style of rt.cuh: every function is called by ALL threads of the block with block-uniform control flow;
// item = thread / 8, lane = thread % 8; all state lives in shared memory (TeCoopState, one per item).
#pragma once
#include "rt.cuh"
#include "te.cuh"

namespace dr {

constexpr uint32_t COOP_LANES = 8;

struct TeCoopPoint {
    Fr c[5];  // X, Y, Z, T, dT: extended coordinates; dT = d * T (only kept for table entries, the second operand of an addition)
};
enum { CX = 0, CY = 1, CZ = 2, CT = 3, CDT = 4 };

struct TeCoopState {
    TeCoopPoint acc;
    Fr m[5];               // products of the first layer
    TeCoopPoint tab[15];   // 1P .. 15P
    TeCoopPoint tab2[15];  // 1Q .. 15Q (two-point Straus only)
    uint32_t k[8];         // scalar of P, raw little-endian limbs
    uint32_t k2[8];        // scalar of Q
    uint32_t live;         // 0: this slot holds no item (lanes idle, phases still run)
    TEExt part[COOP_LANES];  // fixed-base partial sums
};

// Every layer is ONE multiplication executed by all lanes in lockstep: the lanes differ in their operands only (a warp that
// branched per lane would run the four or five products one after another, which is exactly what this file exists to avoid).
DR_HD Fr coop_pick(uint32_t i, const Fr& a0, const Fr& a1, const Fr& a2, const Fr& a3) {
    Fr r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = i == 0 ? a0.v[l] : i == 1 ? a1.v[l] : i == 2 ? a2.v[l] : a3.v[l];
    return r;
}

// first layer of a doubling of `p`: A = X^2, B = Y^2, C = 2 Z^2, E' = (X + Y)^2   (lanes 0..3)
DR_HD void coop_dbl_layer1(uint32_t lane, const TeCoopPoint& p, Fr* m) {
    if (lane > 3) return;
    const Fr own = p.c[lane < 3 ? lane : 0];  // X, Y, Z, X
    const Fr sum = own + p.c[CY];             // only lane 3 uses it: X + Y
    const Fr a = Fr::select(lane == 3, sum, own);
    const Fr sq = a.sqr();
    m[lane] = Fr::select(lane == 2, sq.dbl(), sq);
}
// second layer: dbl-2008-hwcd with a = -5
DR_HD void coop_dbl_layer2(uint32_t lane, const Fr* m, TeCoopPoint& out) {
    if (lane > 3) return;
    const Fr A = m[0], B = m[1], C = m[2];
    const Fr D = fr_mul5(A).neg();
    const Fr E = m[3] - A - B;
    const Fr G = D + B;
    const Fr F = G - C;
    const Fr H = D - B;
    // X = E F, Y = G H, Z = F G, T = E H
    const Fr a = coop_pick(lane, E, G, F, E), b = coop_pick(lane, F, H, G, H);
    out.c[lane] = a * b;
}
// first layer of p + q (q carries dT): A = X1 X2, B = Y1 Y2, C = T1 (d T2), D = Z1 Z2, E' = (X1 + Y1)(X2 + Y2)   (lanes 0..4)
DR_HD void coop_add_layer1(uint32_t lane, const TeCoopPoint& p, const TeCoopPoint& q, Fr* m) {
    if (lane > 4) return;
    // operand coordinates per lane: (X, X) (Y, Y) (T, dT) (Z, Z) (X + Y, X + Y)
    const uint32_t ia = lane == 2 ? CT : lane == 3 ? CZ : lane == 1 ? CY : CX;
    const uint32_t ib = lane == 2 ? CDT : ia;
    Fr a = p.c[ia], b = q.c[ib];
    const Fr sa = a + p.c[CY], sb = b + q.c[CY];  // only lane 4 uses them
    a = Fr::select(lane == 4, sa, a);
    b = Fr::select(lane == 4, sb, b);
    m[lane] = a * b;
}
DR_HD void coop_add_layer2(uint32_t lane, const Fr* m, TeCoopPoint& out) {
    if (lane > 3) return;
    const Fr E = m[4] - m[0] - m[1];
    const Fr F = m[3] - m[2];
    const Fr G = m[3] + m[2];
    const Fr H = m[1] + fr_mul5(m[0]);
    const Fr a = coop_pick(lane, E, G, F, E), b = coop_pick(lane, F, H, G, H);
    out.c[lane] = a * b;
}

// tab[1..14] <- 2T .. 15T from tab[0] = T (any extended representation, Z need not be 1), dT for every entry.  Ends in a sync.
DR_HD void te_table_coop(const BlockCtx& ctx, TeCoopState* st, bool second) {
    DR_THREAD_LOOP(t, ctx) {
        TeCoopState& s = st[t / COOP_LANES];
        TeCoopPoint* tab = second ? s.tab2 : s.tab;
        if (s.live && t % COOP_LANES == 0) tab[0].c[CDT] = tab[0].c[CT] * te_d();
    }
    DR_BLOCK_SYNC();
    for (uint32_t i = 2; i <= 15; i++) {  // 2P = dbl(P), 3P = 2P + P, 4P = dbl(2P), ...
        DR_THREAD_LOOP(t, ctx) {
            TeCoopState& s = st[t / COOP_LANES];
            TeCoopPoint* tab = second ? s.tab2 : s.tab;
            if (s.live) {
                if (i & 1) coop_add_layer1(t % COOP_LANES, tab[i - 2], tab[0], s.m);
                else coop_dbl_layer1(t % COOP_LANES, tab[i / 2 - 1], s.m);
            }
        }
        DR_BLOCK_SYNC();
        DR_THREAD_LOOP(t, ctx) {
            TeCoopState& s = st[t / COOP_LANES];
            TeCoopPoint* tab = second ? s.tab2 : s.tab;
            if (s.live) {
                if (i & 1) coop_add_layer2(t % COOP_LANES, s.m, tab[i - 1]);
                else coop_dbl_layer2(t % COOP_LANES, s.m, tab[i - 1]);
            }
        }
        DR_BLOCK_SYNC();
    }
    DR_THREAD_LOOP(t, ctx) {
        TeCoopState& s = st[t / COOP_LANES];
        TeCoopPoint* tab = second ? s.tab2 : s.tab;
        const uint32_t lane = t % COOP_LANES;
        if (s.live)
            for (uint32_t e = lane ? lane : COOP_LANES; e < 15; e += COOP_LANES) tab[e].c[CDT] = tab[e].c[CT] * te_d();  // entry 0 already has it
    }
    DR_BLOCK_SYNC();
}

// acc <- k * P (+ k2 * Q when nlimbs2 > 0) for every live item of the block: Straus with shared doublings.  Before the call (and a
// block sync): st[item].tab[0] = P (and tab2[0] = Q) in extended coordinates, st[item].k / k2 = the scalars (raw limbs, nlimbs /
// nlimbs2 of them significant, nlimbs >= nlimbs2), st[item].live set.  After the call (which ends in a sync) st[item].acc holds
// X, Y, Z, T.  4-bit fixed windows, most significant first, like te_mul_raw / te_msm_small; the formulas are complete on the
// prime-order subgroup (and for the identity), so leading zero windows simply double the identity.
DR_HD void te_straus_coop(const BlockCtx& ctx, TeCoopState* st, int nlimbs, int nlimbs2) {
    te_table_coop(ctx, st, false);
    if (nlimbs2 > 0) te_table_coop(ctx, st, true);
    DR_THREAD_LOOP(t, ctx) {
        TeCoopState& s = st[t / COOP_LANES];
        if (s.live && t % COOP_LANES == 0) {
            s.acc.c[CX] = Fr::zero();
            s.acc.c[CY] = Fr::one();
            s.acc.c[CZ] = Fr::one();
            s.acc.c[CT] = Fr::zero();
        }
    }
    DR_BLOCK_SYNC();
    for (int w = 8 * nlimbs - 1; w >= 0; w--) {
        if (w != 8 * nlimbs - 1) {
            for (int rep = 0; rep < 4; rep++) {
                DR_THREAD_LOOP(t, ctx) {
                    TeCoopState& s = st[t / COOP_LANES];
                    if (s.live) coop_dbl_layer1(t % COOP_LANES, s.acc, s.m);
                }
                DR_BLOCK_SYNC();
                DR_THREAD_LOOP(t, ctx) {
                    TeCoopState& s = st[t / COOP_LANES];
                    if (s.live) coop_dbl_layer2(t % COOP_LANES, s.m, s.acc);
                }
                DR_BLOCK_SYNC();
            }
        }
        for (int which = 0; which < (w < 8 * nlimbs2 ? 2 : 1); which++) {
            DR_THREAD_LOOP(t, ctx) {
                TeCoopState& s = st[t / COOP_LANES];
                const uint32_t d = ((which ? s.k2 : s.k)[w >> 3] >> (4 * (w & 7))) & 15u;
                if (s.live && d) coop_add_layer1(t % COOP_LANES, s.acc, (which ? s.tab2 : s.tab)[d - 1], s.m);
            }
            DR_BLOCK_SYNC();
            DR_THREAD_LOOP(t, ctx) {
                TeCoopState& s = st[t / COOP_LANES];
                const uint32_t d = ((which ? s.k2 : s.k)[w >> 3] >> (4 * (w & 7))) & 15u;
                if (s.live && d) coop_add_layer2(t % COOP_LANES, s.m, s.acc);
            }
            DR_BLOCK_SYNC();
        }
    }
}
DR_HD void te_mul_coop(const BlockCtx& ctx, TeCoopState* st, int nlimbs) { te_straus_coop(ctx, st, nlimbs, 0); }

// helpers for the callers: load an operand / read the result
DR_HD void coop_set_point(TeCoopPoint& dst, const TEExt& p) {
    dst.c[CX] = p.X;
    dst.c[CY] = p.Y;
    dst.c[CZ] = p.Z;
    dst.c[CT] = p.T;
}
DR_HD TEExt coop_get_point(const TeCoopPoint& src) { return {src.c[CX], src.c[CY], src.c[CZ], src.c[CT]}; }

// part[lane] <- sum over the lane's four windows of the fixed-base table entries selected by k (te_mul_fixed split by windows).
// One phase, no sync inside: the caller syncs, then one lane folds the eight partial sums with te_fold_fixed_coop.
DR_HD void te_mul_fixed_coop_partial(uint32_t lane, const TEPre* tab, const uint32_t* k, TEExt* part) {
    TEExt acc = TEExt::identity();
    constexpr int per = TE_FIXED_WINDOWS / (int)COOP_LANES;
#pragma unroll 1
    for (int w = per * (int)lane; w < per * ((int)lane + 1); w++) {
        uint32_t d = (k[w >> 2] >> (8 * (w & 3))) & 255;
        if (d) acc = te_madd(acc, tab[256 * w + d]);
    }
    part[lane] = acc;
}
DR_HD TEExt te_fold_fixed_coop(const TEExt* part) {
    TEExt acc = part[0];
#pragma unroll 1
    for (uint32_t l = 1; l < COOP_LANES; l++) acc = te_add(acc, part[l]);
    return acc;
}

}  // namespace dr
