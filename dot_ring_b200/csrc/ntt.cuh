// Fr NTT in shared memory, one CTA per transform (n <= 4096 elements).
//
// Replaces dot_ring/ring_proof/polynomial/ntt.pyx:116-163 + bls12_381_scalar.c:333-356
// (bit-reverse gather, log2 n radix-2 DIT rounds, optional scale) behind
// dot_ring/ring_proof/polynomial/fft.py:87-144 (inverse_fft / evaluate_poly_fft).  Input and
// output are in natural order like the reference; the transform is exact field arithmetic, so the
// butterfly schedule is free.
//
// The working set lives in shared memory in limb-planar layout (sm[limb * n + index]) so that
// unit-stride butterflies are bank-conflict free with 32-bit accesses.  Loading and storing go
// through functors, which lets callers fuse the witness-column synthesis, the coset pre-twist of
// the 4x low-degree extension and the 1/n scaling into the transform instead of materialising
// intermediate vectors in HBM.
#pragma once
#include "fp.cuh"
#include "rt.cuh"

namespace dr {

DR_HD uint32_t bit_reverse(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

DR_HD Fr sm_load(const uint32_t* sm, uint32_t n, uint32_t i) {
    Fr r;
#pragma unroll
    for (int l = 0; l < 8; l++) r.v[l] = sm[l * n + i];
    return r;
}
DR_HD void sm_store(uint32_t* sm, uint32_t n, uint32_t i, const Fr& x) {
#pragma unroll
    for (int l = 0; l < 8; l++) sm[l * n + i] = x.v[l];
}

// tw[k] = w^k (Montgomery), k < n/2, for the n-th root of unity w of this transform.
// loader(k): k-th input in natural order.  storer(k, value): k-th output in natural order.
//
// Radix-8 passes: a thread pulls 8 elements that are 2^(s-1) apart into registers, runs the three DIT stages s, s+1, s+2 on
// them (12 butterflies) and writes them back, so the working set crosses shared memory log2(n)/3 times instead of
// log2(n) times (the radix-2 version was limited by the shared-memory instruction queue and by 2-way bank conflicts in
// the first stages: ncu mio_throttle / bank-conflict counters in profiles/).  A final pass takes the remaining 1 or 2 stages.
template <class Loader, class Storer>
DR_HD void ntt_block(const BlockCtx& ctx, uint32_t n, uint32_t logn, const Fr* tw, const Loader& loader, const Storer& storer) {
    uint32_t* sm = (uint32_t*)ctx.smem;
    DR_STRIDE_LOOP(e, n, ctx) { sm_store(sm, n, e, loader(bit_reverse(e, logn))); }
    DR_BLOCK_SYNC();
    uint32_t s = 1;  // next DIT stage (1-based): butterflies of span 2^(s-1)
    while (s <= logn) {
        const uint32_t r = (logn - s + 1 >= 3) ? 3 : (logn - s + 1);  // stages in this pass
        const uint32_t span = 1u << (s - 1);
        const uint32_t group = 1u << r;  // elements per thread-group
        DR_STRIDE_LOOP(gidx, n >> r, ctx) {
            // group index -> (high, j_low): element k of the group sits at high * span * group + k * span + j_low
            const uint32_t j_low = gidx & (span - 1);
            const uint32_t base = ((gidx >> (s - 1)) << (s - 1 + r)) + j_low;
            Fr v[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; k++)
                if (k < group) v[k] = sm_load(sm, n, base + k * span);
#pragma unroll
            for (uint32_t t = 0; t < 3; t++) {
                if (t < r) {
                    const uint32_t hl = 1u << t;  // local span
#pragma unroll
                    for (uint32_t bfly = 0; bfly < 4; bfly++) {
                        if (bfly < (group >> 1)) {
                            const uint32_t lo = ((bfly >> t) << (t + 1)) + (bfly & (hl - 1));
                            // position of the pair inside its stage-(s+t) butterfly block
                            const uint32_t j = (lo & (hl - 1)) * span + j_low;
                            Fr x = v[lo], y = v[lo + hl];
                            if (j) y = y * tw[j << (logn - (s + t))];
                            v[lo] = x + y;
                            v[lo + hl] = x - y;
                        }
                    }
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < 8; k++)
                if (k < group) sm_store(sm, n, base + k * span, v[k]);
        }
        DR_BLOCK_SYNC();
        s += r;
    }
    DR_STRIDE_LOOP(e, n, ctx) { storer(e, sm_load(sm, n, e)); }
}

inline size_t ntt_smem_bytes(uint32_t n) { return (size_t)n * 32; }
// one thread per radix-8 group: n / 8 threads keep every lane busy in the butterfly passes (at least one warp, at most 256)
inline uint32_t ntt_threads(uint32_t n) { return n / 8 < 32 ? 32 : (n / 8 > 256 ? 256 : n / 8); }

// ---- plain batched transform (C ABI dr_fr_ntt; ring-root fixed columns) -------------------------
struct NttPlainBody {
    // grid.x = transform index.  scale may be null.  in/out may alias.
    DR_HD void operator()(const BlockCtx& ctx, const Fr* in, Fr* out, uint32_t n, uint32_t logn, const Fr* tw, const Fr* scale) const {
        const Fr* src = in + (size_t)ctx.bx * n;
        Fr* dst = out + (size_t)ctx.bx * n;
        bool has_scale = scale != nullptr;
        Fr sc = has_scale ? *scale : Fr::one();
        ntt_block(
            ctx, n, logn, tw, [&](uint32_t k) { return src[k]; },
            [&](uint32_t k, const Fr& v) { dst[k] = has_scale ? v * sc : v; });
    }
};

// ---- transforms larger than one CTA's shared memory (n = n1 * n2, both <= 4096): two passes ----------------------
//   X[k1 + n1 k2] = sum_{j2} w_{n2}^{j2 k2} * [ w^{j2 k1} * sum_{j1} x[j1 n2 + j2] w_{n1}^{j1 k1} ]
// Pass 1 (grid n2 x batch): column j2 -> n1-point transform (stride-n2 gather), times w^{j2 k1}, into tmp[j2 n1 + k1].
// Pass 2 (grid n1 x batch): row k1 -> n2-point transform over tmp[j2 n1 + k1], natural-order scatter to out[k1 + n1 k2].
// Every element is a full 32-byte sector, so the strided accesses cost no sector efficiency; HBM traffic is two reads and
// two writes per element.  wfull[i] = w^i for i < n; tw1 / tw2 are the half-size twiddle tables of the two sub-transforms.
struct NttLargePass1Body {
    DR_HD void operator()(const BlockCtx& ctx, const Fr* in, Fr* tmp, uint32_t n1, uint32_t logn1, uint32_t n2, const Fr* tw1, const Fr* wfull) const {
        const size_t n = (size_t)n1 * n2;
        const Fr* src = in + (size_t)ctx.by * n;
        Fr* dst = tmp + (size_t)ctx.by * n;
        const uint32_t j2 = ctx.bx;
        ntt_block(
            ctx, n1, logn1, tw1, [&](uint32_t k) { return src[(size_t)k * n2 + j2]; },
            [&](uint32_t k1, const Fr& v) { dst[(size_t)j2 * n1 + k1] = (j2 && k1) ? v * wfull[(size_t)j2 * k1] : v; });
    }
};
struct NttLargePass2Body {
    DR_HD void operator()(const BlockCtx& ctx, const Fr* tmp, Fr* out, uint32_t n1, uint32_t n2, uint32_t logn2, const Fr* tw2, const Fr* scale) const {
        const size_t n = (size_t)n1 * n2;
        const Fr* src = tmp + (size_t)ctx.by * n;
        Fr* dst = out + (size_t)ctx.by * n;
        const uint32_t k1 = ctx.bx;
        bool has_scale = scale != nullptr;
        Fr sc = has_scale ? *scale : Fr::one();
        ntt_block(
            ctx, n2, logn2, tw2, [&](uint32_t j2) { return src[(size_t)j2 * n1 + k1]; },
            [&](uint32_t k2, const Fr& v) { dst[k1 + (size_t)n1 * k2] = has_scale ? v * sc : v; });
    }
};

// canonical little-endian bytes <-> Montgomery limbs, elementwise
struct FrToMontBody {
    DR_HD void operator()(const BlockCtx& ctx, const uint8_t* in, Fr* out, size_t count, uint32_t* bad_flag) const {
        DR_THREAD_LOOP(t, ctx) {
            size_t i = (size_t)ctx.bx * ctx.nthreads + t;
            if (i < count) {
                Fr r;
                fr_from_le_bytes_raw(r, in + 32 * i);
                if (!r.is_canonical_raw()) {
                    // reduce once or more: value < 2^256 < 3r
                    if (bad_flag) *bad_flag = 1;
                    while (!r.is_canonical_raw()) Fr::sub_mod_inplace(r.v);
                }
                out[i] = r.to_mont();
            }
        }
    }
};
struct FrFromMontBody {
    DR_HD void operator()(const BlockCtx& ctx, const Fr* in, uint8_t* out, size_t count) const {
        DR_THREAD_LOOP(t, ctx) {
            size_t i = (size_t)ctx.bx * ctx.nthreads + t;
            if (i < count) fr_to_le_bytes_raw(out + 32 * i, in[i].from_mont());
        }
    }
};

}  // namespace dr
