// Montgomery prime-field arithmetic on 32-bit limbs for BLS12-381 Fq (12 limbs) and Fr (8 limbs).
//
// Replaces the reference's 4x64-bit CIOS C code (dot_ring/curve/native_field/bls12_381_scalar.c:
// 43-257, constants :7-27) and blst's Fq assembly.  On the device every operation is one inline PTX
// block generated (and verified against big integers) by tools/gen_field_asm.py; on the host the
// same functions fall back to portable uint64 arithmetic, which is what the CPU-side unit tests and
// the few host-only paths (pairing, setup) use.
//
// Elements are little-endian limb vectors in Montgomery form (x * 2^(32 N) mod p), always fully
// reduced (< p).
#pragma once
#include <cstdint>
#include <cstring>

#include "gen/field_consts.h"
#include "gen/field_asm.inc"

#if defined(__CUDACC__)
#define DR_HD __host__ __device__ __forceinline__
#define DR_D __device__ __forceinline__
// large, rarely executed helpers (inversions, exponentiations, exceptional group-law cases) stay out of
// line on the device: it keeps the hot loops inside the instruction cache and the build time sane
#define DR_HD_COLD inline __host__ __device__ __noinline__
#else
#define DR_HD inline
#define DR_D inline
#define DR_HD_COLD inline
#endif

namespace dr {

// Out-of-line multiplication.  A 12-limb Montgomery multiplication is ~6 KB of straight-line SASS; a mixed G1
// addition inlines ten of them, which overflows the instruction cache (ncu: `no_instruction` was the second
// largest stall of the commit kernel).  Operands and result travel by value, which the device ABI keeps
// entirely in registers (no local-memory traffic), so a call costs a CALL/RET pair and a few moves.
// -DDR_FQ_MUL_INLINE / -DDR_FR_MUL_INLINE restore full inlining for A/B measurements.
#if defined(__CUDA_ARCH__)
struct Limbs12 {
    uint32_t v[12];
};
struct Limbs8 {
    uint32_t v[8];
};
static __device__ __noinline__ Limbs12 fq_mul_outofline(Limbs12 a, Limbs12 b) {
    Limbs12 r;
    fq_mul_ptx(r.v, a.v, b.v);
    return r;
}
static __device__ __noinline__ Limbs12 fq_sqr_outofline(Limbs12 a) {
    Limbs12 r;
    fq_sqr_ptx(r.v, a.v);
    return r;
}
static __device__ __noinline__ Limbs8 fr_mul_outofline(Limbs8 a, Limbs8 b) {
    Limbs8 r;
    fr_mul_ptx(r.v, a.v, b.v);
    return r;
}
static __device__ __noinline__ Limbs8 fr_sqr_outofline(Limbs8 a) {
    Limbs8 r;
    fr_sqr_ptx(r.v, a.v);
    return r;
}
#endif

struct FqTag {
    static constexpr int N = DR_FQ_LIMBS;
    static constexpr uint32_t M0 = DR_FQ_M0;
    DR_HD static uint32_t mod(int i) {
        constexpr uint32_t m[N] = DR_FQ_MOD;
        return m[i];
    }
    DR_HD static uint32_t r1(int i) {
        constexpr uint32_t m[N] = DR_FQ_R1;
        return m[i];
    }
    DR_HD static uint32_t r2(int i) {
        constexpr uint32_t m[N] = DR_FQ_R2;
        return m[i];
    }
#if defined(__CUDA_ARCH__)
#if defined(DR_FQ_MUL_INLINE)
    DR_D static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { fq_mul_ptx(r, a, b); }
    DR_D static void sqr(uint32_t* r, const uint32_t* a) { fq_sqr_ptx(r, a); }
#else
    DR_D static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        Limbs12 x, y;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            x.v[i] = a[i];
            y.v[i] = b[i];
        }
        Limbs12 z = fq_mul_outofline(x, y);
#pragma unroll
        for (int i = 0; i < 12; i++) r[i] = z.v[i];
    }
    DR_D static void sqr(uint32_t* r, const uint32_t* a) {
        Limbs12 x;
#pragma unroll
        for (int i = 0; i < 12; i++) x.v[i] = a[i];
        Limbs12 z = fq_sqr_outofline(x);
#pragma unroll
        for (int i = 0; i < 12; i++) r[i] = z.v[i];
    }
#endif
    DR_D static void add(uint32_t* r, const uint32_t* a, const uint32_t* b) { fq_add_ptx(r, a, b); }
    DR_D static void sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { fq_sub_ptx(r, a, b); }
#endif
};

struct FrTag {
    static constexpr int N = DR_FR_LIMBS;
    static constexpr uint32_t M0 = DR_FR_M0;
    DR_HD static uint32_t mod(int i) {
        constexpr uint32_t m[N] = DR_FR_MOD;
        return m[i];
    }
    DR_HD static uint32_t r1(int i) {
        constexpr uint32_t m[N] = DR_FR_R1;
        return m[i];
    }
    DR_HD static uint32_t r2(int i) {
        constexpr uint32_t m[N] = DR_FR_R2;
        return m[i];
    }
#if defined(__CUDA_ARCH__)
#if !defined(DR_FR_MUL_CALL)
    DR_D static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { fr_mul_ptx(r, a, b); }
    DR_D static void sqr(uint32_t* r, const uint32_t* a) { fr_sqr_ptx(r, a); }
#else
    DR_D static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        Limbs8 x, y;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x.v[i] = a[i];
            y.v[i] = b[i];
        }
        Limbs8 z = fr_mul_outofline(x, y);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = z.v[i];
    }
    DR_D static void sqr(uint32_t* r, const uint32_t* a) {
        Limbs8 x;
#pragma unroll
        for (int i = 0; i < 8; i++) x.v[i] = a[i];
        Limbs8 z = fr_sqr_outofline(x);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = z.v[i];
    }
#endif
    DR_D static void add(uint32_t* r, const uint32_t* a, const uint32_t* b) { fr_add_ptx(r, a, b); }
    DR_D static void sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { fr_sub_ptx(r, a, b); }
#endif
};

struct FnTag {  // Bandersnatch prime-subgroup order (VRF scalar arithmetic)
    static constexpr int N = DR_FN_LIMBS;
    static constexpr uint32_t M0 = DR_FN_M0;
    DR_HD static uint32_t mod(int i) {
        constexpr uint32_t m[N] = DR_FN_MOD;
        return m[i];
    }
    DR_HD static uint32_t r1(int i) {
        constexpr uint32_t m[N] = DR_FN_R1;
        return m[i];
    }
    DR_HD static uint32_t r2(int i) {
        constexpr uint32_t m[N] = DR_FN_R2;
        return m[i];
    }
#if defined(__CUDA_ARCH__)
    DR_D static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { fn_mul_ptx(r, a, b); }
    DR_D static void sqr(uint32_t* r, const uint32_t* a) { fn_sqr_ptx(r, a); }
    DR_D static void add(uint32_t* r, const uint32_t* a, const uint32_t* b) { fn_add_ptx(r, a, b); }
    DR_D static void sub(uint32_t* r, const uint32_t* a, const uint32_t* b) { fn_sub_ptx(r, a, b); }
#endif
};

template <class T>
struct alignas(16) Fp {
    static constexpr int N = T::N;
    uint32_t v[N];

    DR_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    DR_HD static Fp one() {  // Montgomery form of 1
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = T::r1(i);
        return r;
    }
    DR_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = T::r2(i);
        return r;
    }
    DR_HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = T::mod(i);
        return r;
    }
    DR_HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i];
        return acc == 0;
    }
    DR_HD bool operator==(const Fp& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i] ^ o.v[i];
        return acc == 0;
    }
    DR_HD bool operator!=(const Fp& o) const { return !(*this == o); }

    // ---- portable limb helpers (host path, also usable on device for debugging) ----
    DR_HD static bool geq_mod(const uint32_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            uint32_t m = T::mod(i);
            if (a[i] != m) return a[i] > m;
        }
        return true;
    }
    DR_HD static void sub_mod_inplace(uint32_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            uint64_t t = (uint64_t)a[i] - T::mod(i) - borrow;
            a[i] = (uint32_t)t;
            borrow = (t >> 32) & 1;
        }
    }
    DR_HD static void mul_portable(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        uint32_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            uint64_t c = 0;
            for (int j = 0; j < N; j++) {
                uint64_t x = (uint64_t)a[j] * b[i] + t[j] + c;
                t[j] = (uint32_t)x;
                c = x >> 32;
            }
            uint64_t x = (uint64_t)t[N] + c;
            t[N] = (uint32_t)x;
            t[N + 1] = (uint32_t)(x >> 32);
            uint32_t m = t[0] * T::M0;
            c = ((uint64_t)m * T::mod(0) + t[0]) >> 32;
            for (int j = 1; j < N; j++) {
                uint64_t y = (uint64_t)m * T::mod(j) + t[j] + c;
                t[j - 1] = (uint32_t)y;
                c = y >> 32;
            }
            x = (uint64_t)t[N] + c;
            t[N - 1] = (uint32_t)x;
            t[N] = t[N + 1] + (uint32_t)(x >> 32);
        }
        if (t[N] || geq_mod(t)) sub_mod_inplace(t);
        for (int i = 0; i < N; i++) r[i] = t[i];
    }
    DR_HD static void add_portable(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        uint32_t t[N];
        uint64_t c = 0;
        for (int i = 0; i < N; i++) {
            uint64_t x = (uint64_t)a[i] + b[i] + c;
            t[i] = (uint32_t)x;
            c = x >> 32;
        }
        if (c || geq_mod(t)) sub_mod_inplace(t);
        for (int i = 0; i < N; i++) r[i] = t[i];
    }
    DR_HD static void sub_portable(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        uint32_t t[N];
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            uint64_t x = (uint64_t)a[i] - b[i] - borrow;
            t[i] = (uint32_t)x;
            borrow = (x >> 32) & 1;
        }
        if (borrow) {
            uint64_t c = 0;
            for (int i = 0; i < N; i++) {
                uint64_t x = (uint64_t)t[i] + T::mod(i) + c;
                t[i] = (uint32_t)x;
                c = x >> 32;
            }
        }
        for (int i = 0; i < N; i++) r[i] = t[i];
    }

    // ---- field operations ----
    DR_HD friend Fp operator*(const Fp& a, const Fp& b) {
        Fp r;
#if defined(__CUDA_ARCH__) && !defined(DR_PORTABLE_FIELD)
        T::mul(r.v, a.v, b.v);
#else
        mul_portable(r.v, a.v, b.v);
#endif
        return r;
    }
    DR_HD Fp sqr() const {
        Fp r;
#if defined(__CUDA_ARCH__) && !defined(DR_PORTABLE_FIELD)
        T::sqr(r.v, v);
#else
        mul_portable(r.v, v, v);
#endif
        return r;
    }
    DR_HD friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
#if defined(__CUDA_ARCH__) && !defined(DR_PORTABLE_FIELD)
        T::add(r.v, a.v, b.v);
#else
        add_portable(r.v, a.v, b.v);
#endif
        return r;
    }
    DR_HD friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
#if defined(__CUDA_ARCH__) && !defined(DR_PORTABLE_FIELD)
        T::sub(r.v, a.v, b.v);
#else
        sub_portable(r.v, a.v, b.v);
#endif
        return r;
    }
    DR_HD Fp neg() const { return zero() - *this; }
    DR_HD Fp dbl() const { return *this + *this; }
    DR_HD Fp& operator*=(const Fp& o) { return *this = *this * o; }
    DR_HD Fp& operator+=(const Fp& o) { return *this = *this + o; }
    DR_HD Fp& operator-=(const Fp& o) { return *this = *this - o; }

    // Montgomery <-> canonical
    DR_HD Fp to_mont() const { return *this * r2(); }
    DR_HD Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }
    // canonical little-endian limbs must already be < p
    DR_HD bool is_canonical_raw() const { return !geq_mod(v); }

    // x^e for a little-endian limb exponent (variable time; exponents here are public)
    DR_HD_COLD Fp pow(const uint32_t* e, int nlimbs) const {
        Fp acc = one();
        bool started = false;
#pragma unroll 1
        for (int i = nlimbs - 1; i >= 0; i--) {
#pragma unroll 1
            for (int b = 31; b >= 0; b--) {
                if (started) acc = acc.sqr();
                if ((e[i] >> b) & 1) {
                    acc = started ? acc * *this : *this;
                    started = true;
                }
            }
        }
        return acc;
    }
    // Fermat inverse x^(p-2); 0 -> 0.  Kept for A/B runs and as the cross-check of inv() in the unit tests.
    DR_HD_COLD Fp inv_fermat() const {
        uint32_t e[N];
#pragma unroll
        for (int i = 0; i < N; i++) e[i] = T::mod(i);
        uint32_t borrow = 2;  // p - 2 with borrow propagation (Fr's low limb is 1)
        for (int i = 0; i < N && borrow; i++) {
            uint32_t old = e[i];
            e[i] = old - borrow;
            borrow = old < borrow ? 1 : 0;
        }
        return pow(e, N);
    }
    // Modular inverse by Bernstein-Yang division steps ("safegcd"), batched 30 at a time.  A batch looks only at the low words of
    // f and g: 30 branch-free steps on two 32-bit integers produce a 2 x 2 transition matrix with entries below 2^30, which is then
    // applied to the full-length f, g (exact division by 2^30) and to the Bezout pair d, e (division by 2^30 modulo p, using
    // p^-1 mod 2^30).  Against the bit-at-a-time binary algorithm below this replaces ~1.4 bits(p) passes over N-limb integers by
    // ~bits(p) / 21 of them plus cheap word steps: measured on B200 as ONE dependent chain (the regime of the latency-bound
    // kernels) an Fr inversion drops from 94 us (Kaliski; 179 us Fermat) to the figure in profiles/, an Fq inversion from 184 us.
    // Signed 30-bit limbs, ranges and the update formulas follow the published algorithm (Bernstein & Yang 2019; the 30-bit
    // batching is the layout libsecp256k1 documents in doc/safegcd_implementation.md).  The number of batches depends on the
    // operand, like gmpy2's invert in the reference.  0 -> 0.
    static constexpr int L30 = (32 * N + 29) / 30;
    DR_HD_COLD Fp inv() const {
        if (is_zero()) return zero();
        const int32_t M30 = (int32_t)(0xFFFFFFFFu >> 2);
        int32_t f[L30], g[L30], d[L30], e[L30], md30[L30];
        // 32-bit limbs -> 30-bit limbs
        auto to30 = [&](auto limb, int32_t* out) {
#pragma unroll
            for (int i = 0; i < L30; i++) {
                const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
                uint64_t lo = w < N ? limb(w) : 0u, hi = w + 1 < N ? limb(w + 1) : 0u;
                out[i] = (int32_t)(uint32_t)(((lo | (hi << 32)) >> sh) & (uint64_t)M30);
            }
        };
        to30([](int i) { return (uint64_t)T::mod(i); }, f);
        to30([](int i) { return (uint64_t)T::mod(i); }, md30);
        to30([this](int i) { return (uint64_t)v[i]; }, g);
#pragma unroll
        for (int i = 0; i < L30; i++) d[i] = e[i] = 0;
        e[0] = 1;
        const uint32_t pinv30 = (0u - T::M0) & (uint32_t)M30;  // p^-1 mod 2^30 (M0 = -p^-1 mod 2^32)
        int32_t zeta = -1;                                      // -(delta + 1/2), delta = 1/2
#pragma unroll 1
        for (int batch = 0; batch < (49 * 32 * N + 57) / 17 / 30 + 2; batch++) {
            // 30 division steps on the low words -> matrix (u v; q r)
            uint32_t u = 1, vv = 0, q = 0, r = 1;
            uint32_t fl = (uint32_t)f[0], gl = (uint32_t)g[0];
#pragma unroll 6
            for (int i = 0; i < 30; i++) {
                uint32_t c1 = (uint32_t)(zeta >> 31), c2 = 0u - (gl & 1u);
                uint32_t x = (fl ^ c1) - c1, y = (u ^ c1) - c1, z = (vv ^ c1) - c1;
                gl += x & c2;
                q += y & c2;
                r += z & c2;
                c1 &= c2;
                zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
                fl += gl & c1;
                u += q & c1;
                vv += r & c1;
                gl >>= 1;
                u <<= 1;
                vv <<= 1;
            }
            const int32_t mu = (int32_t)u, mv = (int32_t)vv, mq = (int32_t)q, mr = (int32_t)r;
            // (d, e) <- (u d + v e, q d + r e) / 2^30 mod p, kept in (-2p, p)
            {
                const int32_t sd = d[L30 - 1] >> 31, se = e[L30 - 1] >> 31;
                int32_t md = (mu & sd) + (mv & se), me = (mq & sd) + (mr & se);
                int64_t cd = (int64_t)mu * d[0] + (int64_t)mv * e[0], ce = (int64_t)mq * d[0] + (int64_t)mr * e[0];
                md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
                me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
                cd += (int64_t)md30[0] * md;
                ce += (int64_t)md30[0] * me;
                cd >>= 30;
                ce >>= 30;
#pragma unroll
                for (int i = 1; i < L30; i++) {
                    cd += (int64_t)mu * d[i] + (int64_t)mv * e[i] + (int64_t)md30[i] * md;
                    ce += (int64_t)mq * d[i] + (int64_t)mr * e[i] + (int64_t)md30[i] * me;
                    d[i - 1] = (int32_t)cd & M30;
                    e[i - 1] = (int32_t)ce & M30;
                    cd >>= 30;
                    ce >>= 30;
                }
                d[L30 - 1] = (int32_t)cd;
                e[L30 - 1] = (int32_t)ce;
            }
            // (f, g) <- (u f + v g, q f + r g) / 2^30 (exact)
            {
                int64_t cf = (int64_t)mu * f[0] + (int64_t)mv * g[0], cg = (int64_t)mq * f[0] + (int64_t)mr * g[0];
                cf >>= 30;
                cg >>= 30;
                int32_t gnz = 0;
#pragma unroll
                for (int i = 1; i < L30; i++) {
                    cf += (int64_t)mu * f[i] + (int64_t)mv * g[i];
                    cg += (int64_t)mq * f[i] + (int64_t)mr * g[i];
                    f[i - 1] = (int32_t)cf & M30;
                    g[i - 1] = (int32_t)cg & M30;
                    gnz |= g[i - 1];
                    cf >>= 30;
                    cg >>= 30;
                }
                f[L30 - 1] = (int32_t)cf;
                g[L30 - 1] = (int32_t)cg;
                if (!(gnz | g[L30 - 1])) break;  // g == 0: f = +-gcd = +-1 and d = +-x^-1
            }
        }
        // normalise d from (-2p, p) to [0, p), negated when f = -1
        {
            const int32_t sign = f[L30 - 1] >> 31;
            int32_t add = d[L30 - 1] >> 31;
#pragma unroll
            for (int i = 0; i < L30; i++) d[i] = ((d[i] + (md30[i] & add)) ^ sign) - sign;
#pragma unroll
            for (int i = 0; i + 1 < L30; i++) {
                d[i + 1] += d[i] >> 30;
                d[i] &= M30;
            }
            add = d[L30 - 1] >> 31;
#pragma unroll
            for (int i = 0; i < L30; i++) d[i] += md30[i] & add;
#pragma unroll
            for (int i = 0; i + 1 < L30; i++) {
                d[i + 1] += d[i] >> 30;
                d[i] &= M30;
            }
        }
        // 30-bit limbs -> 32-bit limbs
        Fp y;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int bit = 32 * i, w = bit / 30, sh = bit % 30;
            uint64_t acc = (uint64_t)(uint32_t)d[w] >> sh;
            if (w + 1 < L30) acc |= (uint64_t)(uint32_t)d[w + 1] << (30 - sh);
            if (w + 2 < L30) acc |= (uint64_t)(uint32_t)d[w + 2] << (60 - sh);
            y.v[i] = (uint32_t)acc;
        }
        // y = (x R)^-1 as a plain residue; the Montgomery form of x^-1 is y R^2: two products with R^2 (each contributes R)
        return (y * r2()) * r2();
    }
    // Kaliski's binary "Montgomery inverse" (shift / subtract steps on N-limb integers, no multiplications; the previous inv()):
    // phase 1 turns (p, x) into x^-1 * 2^k with bits(p) <= k <= 2 bits(p); phase 2 multiplies the power of two away.  About
    // 1.4 bits(p) branch-free iterations of ~13 N word operations: ~3x fewer issue slots than the 1.5 bits(p) Montgomery
    // multiplications of Fermat's method and, where a kernel is latency-bound (one thread per proof: every multiplication is one
    // long carry chain), more than 10x shorter.  The iteration count depends on the operand (as gmpy2's invert does in the
    // reference); all data paths inside an iteration are selects.  0 -> 0.
    DR_HD_COLD Fp inv_kaliski() const {
        if (is_zero()) return zero();
        uint32_t u[N], w[N], r[N], s[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            u[i] = T::mod(i);
            w[i] = v[i];
            r[i] = 0;
            s[i] = 0;
        }
        s[0] = 1;
        uint32_t k = 0;
#pragma unroll 1
        for (;;) {
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) nz |= w[i];
            if (!nz) break;
            // d = u - w, borrow <=> u < w
            uint32_t d[N];
            uint32_t borrow = 0, dnz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t t = (uint64_t)u[i] - w[i] - borrow;
                d[i] = (uint32_t)t;
                borrow = (uint32_t)(t >> 32) & 1u;
                dnz |= d[i];
            }
            const bool u_odd = u[0] & 1u, w_odd = w[0] & 1u;
            const bool u_gt = !borrow && dnz;
            const bool sub = u_odd && w_odd;                // both odd: the larger one absorbs the difference
            const bool swap = u_odd && (!w_odd || !u_gt);   // the step acts on (w, s) instead of (u, r)
            // x <- (x - [sub] y) >> 1 ; pp <- pp + [sub] qq ; qq <- qq << 1      with (x, y, pp, qq) = swap ? (w, u, s, r) : (u, w, r, s)
            uint32_t x[N], pp[N], qq[N];
            uint32_t ncarry = 1;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t neg = (uint64_t)(~d[i]) + ncarry;  // -d = w - u
                ncarry = (uint32_t)(neg >> 32);
                uint32_t diff = swap ? (uint32_t)neg : d[i];
                x[i] = sub ? diff : (swap ? w[i] : u[i]);
                pp[i] = swap ? s[i] : r[i];
                qq[i] = swap ? r[i] : s[i];
            }
            uint32_t carry = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t t = (uint64_t)pp[i] + (sub ? qq[i] : 0u) + carry;
                pp[i] = (uint32_t)t;
                carry = (uint32_t)(t >> 32);
            }
#pragma unroll
            for (int i = 0; i < N; i++) x[i] = (x[i] >> 1) | (i + 1 < N ? x[i + 1] << 31 : 0u);
#pragma unroll
            for (int i = N - 1; i >= 0; i--) qq[i] = (qq[i] << 1) | (i > 0 ? qq[i - 1] >> 31 : 0u);
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (swap) {
                    w[i] = x[i];
                    s[i] = pp[i];
                    r[i] = qq[i];
                } else {
                    u[i] = x[i];
                    r[i] = pp[i];
                    s[i] = qq[i];
                }
            }
            k++;
        }
        // r < 2p: reduce, then negate:  x^-1 * 2^k = p - r
        if (geq_mod(r)) sub_mod_inplace(r);
        Fp y;
        {
            uint32_t borrow = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t t = (uint64_t)T::mod(i) - r[i] - borrow;
                y.v[i] = (uint32_t)t;
                borrow = (uint32_t)(t >> 32) & 1u;
            }
        }
        // The operand was x R (Montgomery form), so y = x^-1 R^-1 2^k and the Montgomery form of the inverse is y * 2^(64 N - k):
        // every product with r2() contributes 2^(32 N), the last factor is the single bit 2^f as a raw integer (2^f < R; the
        // reduction only needs one operand below p).
        uint32_t f = 64u * N - k;
#pragma unroll 1
        while (f >= 32u * N) {
            y = y * r2();
            f -= 32u * N;
        }
        y = y * r2();
        Fp bit = zero();
        bit.v[f >> 5] = 1u << (f & 31);
        return y * bit;
    }
    // Quadratic character of the element: 1 (non-zero square), -1 (non-square), 0 (zero), by the binary Jacobi algorithm (shifts and
    // subtractions on N-limb integers, branch-free steps like inv()).  The Montgomery factor 2^(32 N) is an even power of two, a
    // square, so the symbol is taken straight from the Montgomery limbs.
    DR_HD_COLD int legendre() const {
        uint32_t a[N], n[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            a[i] = v[i];
            n[i] = T::mod(i);
        }
        uint32_t flip = 0;  // parity of the sign changes
#pragma unroll 1
        for (;;) {
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) nz |= a[i];
            if (!nz) break;
            const bool odd = a[0] & 1u;
            // d = a - n (borrow <=> a < n)
            uint32_t d[N], borrow = 0;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t t = (uint64_t)a[i] - n[i] - borrow;
                d[i] = (uint32_t)t;
                borrow = (uint32_t)(t >> 32) & 1u;
            }
            const bool swap = odd && borrow;  // a < n, both odd: (a / n) = (n / a) * (-1)^((a-1)(n-1)/4)
            if (swap) flip ^= ((a[0] & 3u) == 3u && (n[0] & 3u) == 3u) ? 1u : 0u;
            uint32_t ncarry = 1;
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint64_t neg = (uint64_t)(~d[i]) + ncarry;  // n - a
                ncarry = (uint32_t)(neg >> 32);
                const uint32_t big = swap ? (uint32_t)neg : d[i];  // |a - n|
                const uint32_t small = swap ? a[i] : n[i];
                a[i] = odd ? big : a[i];
                n[i] = odd ? small : n[i];
            }
            // a is even now (it was, or it is a difference of two odd numbers): halve, (2 / n) = -1 iff n = 3, 5 mod 8
            const uint32_t r = n[0] & 7u;
            flip ^= (r == 3u || r == 5u) ? 1u : 0u;
#pragma unroll
            for (int i = 0; i < N; i++) a[i] = (a[i] >> 1) | (i + 1 < N ? a[i + 1] << 31 : 0u);
        }
        uint32_t rest = n[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) rest |= n[i];
        if (rest) return 0;  // gcd != 1: the element is zero (the modulus is prime)
        return flip ? -1 : 1;
    }
    DR_HD static Fp from_u32(uint32_t x) {
        Fp r = zero();
        r.v[0] = x;
        return r.to_mont();
    }
    // conditional select (branch-free): c ? a : b
    DR_HD static Fp select(bool c, const Fp& a, const Fp& b) {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = c ? a.v[i] : b.v[i];
        return r;
    }
};

typedef Fp<FqTag> Fq;
typedef Fp<FrTag> Fr;
typedef Fp<FnTag> Fn;

// Reduce an arbitrary little-endian byte string (len <= 64) into the field, Montgomery form:
// Horner over 16-byte chunks so every partial operand stays canonical.
template <class F>
DR_HD F fp_from_le_bytes_mod(const uint8_t* in, int len) {
    static_assert(F::N == 8, "8-limb fields only");
    F two128 = F::zero();
    two128.v[4] = 1;
    two128 = two128.to_mont();
    F acc = F::zero();
    int nchunks = (len + 15) / 16;
#pragma unroll 1
    for (int c = nchunks - 1; c >= 0; c--) {
        F chunk = F::zero();
        for (int b = 0; b < 16; b++) {
            int idx = 16 * c + b;
            if (idx < len) chunk.v[b >> 2] |= (uint32_t)in[idx] << (8 * (b & 3));
        }
        acc = acc * two128 + chunk.to_mont();
    }
    return acc;
}

// ---- byte codecs used at the C ABI ---------------------------------------------------------
// Fr: 32-byte little-endian canonical.  Fq: 48-byte big-endian canonical (zcash).
DR_HD void fr_from_le_bytes_raw(Fr& out, const uint8_t* in) {
#pragma unroll
    for (int i = 0; i < 8; i++)
        out.v[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24);
}
DR_HD void fr_to_le_bytes_raw(uint8_t* out, const Fr& in) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)in.v[i];
        out[4 * i + 1] = (uint8_t)(in.v[i] >> 8);
        out[4 * i + 2] = (uint8_t)(in.v[i] >> 16);
        out[4 * i + 3] = (uint8_t)(in.v[i] >> 24);
    }
}
DR_HD void fq_from_be_bytes_raw(Fq& out, const uint8_t* in) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint8_t* p = in + 44 - 4 * i;
        out.v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
}
DR_HD void fq_to_be_bytes_raw(uint8_t* out, const Fq& in) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint8_t* p = out + 44 - 4 * i;
        p[0] = (uint8_t)(in.v[i] >> 24);
        p[1] = (uint8_t)(in.v[i] >> 16);
        p[2] = (uint8_t)(in.v[i] >> 8);
        p[3] = (uint8_t)in.v[i];
    }
}

}  // namespace dr
