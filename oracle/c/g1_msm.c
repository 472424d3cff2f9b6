/* G1 multi-scalar multiplication for the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of what the reference obtains from the third-party blst library through
 *   dot_ring/ring_proof/pcs/kzg.py:147-149,170-173  (blst.P1_Affines.mult_pippenger over SRS memory)
 * i.e. sum_i k_i * P_i over BLS12-381 G1, computed with the published bucket ("Pippenger") method.
 * blst itself is absent from /root/reference (cloned un-pinned at build time, setup.py:24-25), so this
 * follows the textbook algorithm: 6x64-bit Montgomery field arithmetic, Jacobian coordinates, unsigned
 * c-bit windows.  Loaded by oracle/bls12_381.py via ctypes when built (oracle/c/Makefile); the oracle falls
 * back to its pure-Python MSM otherwise.  Never linked into the product library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[6]; } fq;
static const fq MOD = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const fq ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};   /* 2^384 mod p */
static const fq RR = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL, 0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};    /* 2^768 mod p */
static const uint64_t NINV = 0x89f3fffcfffcfffdULL; /* -p^-1 mod 2^64 */

static int fq_is_zero(const fq* a) { uint64_t t = 0; for (int i = 0; i < 6; i++) t |= a->v[i]; return t == 0; }
static int fq_eq(const fq* a, const fq* b) { uint64_t t = 0; for (int i = 0; i < 6; i++) t |= a->v[i] ^ b->v[i]; return t == 0; }
static int geq_mod(const uint64_t* a) {
    for (int i = 5; i >= 0; i--) { if (a[i] != MOD.v[i]) return a[i] > MOD.v[i]; }
    return 1;
}
static void sub_mod(uint64_t* a) {
    uint64_t borrow = 0;
    for (int i = 0; i < 6; i++) { u128 t = (u128)a[i] - MOD.v[i] - borrow; a[i] = (uint64_t)t; borrow = (uint64_t)(t >> 64) & 1; }
}
static void fq_add(fq* r, const fq* a, const fq* b) {
    uint64_t t[6]; uint64_t c = 0;
    for (int i = 0; i < 6; i++) { u128 s = (u128)a->v[i] + b->v[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    if (c || geq_mod(t)) sub_mod(t);
    memcpy(r->v, t, 48);
}
static void fq_sub(fq* r, const fq* a, const fq* b) {
    uint64_t t[6]; uint64_t borrow = 0;
    for (int i = 0; i < 6; i++) { u128 s = (u128)a->v[i] - b->v[i] - borrow; t[i] = (uint64_t)s; borrow = (uint64_t)(s >> 64) & 1; }
    if (borrow) { uint64_t c = 0; for (int i = 0; i < 6; i++) { u128 s = (u128)t[i] + MOD.v[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } }
    memcpy(r->v, t, 48);
}
static void fq_mul(fq* r, const fq* a, const fq* b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 6; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 6; j++) { u128 s = (u128)a->v[j] * b->v[i] + t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        u128 s = (u128)t[6] + c; t[6] = (uint64_t)s; t[7] = (uint64_t)(s >> 64);
        uint64_t m = t[0] * NINV;
        c = (uint64_t)(((u128)m * MOD.v[0] + t[0]) >> 64);
        for (int j = 1; j < 6; j++) { u128 s2 = (u128)m * MOD.v[j] + t[j] + c; t[j - 1] = (uint64_t)s2; c = (uint64_t)(s2 >> 64); }
        s = (u128)t[6] + c; t[5] = (uint64_t)s; t[6] = t[7] + (uint64_t)(s >> 64);
    }
    if (t[6] || geq_mod(t)) sub_mod(t);
    memcpy(r->v, t, 48);
}
static void fq_sqr(fq* r, const fq* a) { fq_mul(r, a, a); }
static void fq_from_be(fq* r, const uint8_t* b) {
    fq raw;
    for (int i = 0; i < 6; i++) { uint64_t w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[8 * (5 - i) + k]; raw.v[i] = w; }
    fq_mul(r, &raw, &RR);
}
static void fq_to_be(uint8_t* b, const fq* a) {
    fq one_raw = {{1, 0, 0, 0, 0, 0}}, raw;
    fq_mul(&raw, a, &one_raw);
    for (int i = 0; i < 6; i++) for (int k = 0; k < 8; k++) b[8 * (5 - i) + k] = (uint8_t)(raw.v[i] >> (56 - 8 * k));
}
static void fq_inv(fq* r, const fq* a) { /* a^(p-2) */
    uint64_t e[6]; memcpy(e, MOD.v, 48); e[0] -= 2;
    fq acc = ONE;
    for (int i = 5; i >= 0; i--) for (int b = 63; b >= 0; b--) { fq_sqr(&acc, &acc); if ((e[i] >> b) & 1) fq_mul(&acc, &acc, a); }
    *r = acc;
}

typedef struct { fq x, y, z; } jac; /* z == 0 <=> infinity */
typedef struct { fq x, y; } aff;

static void jac_dbl(jac* r, const jac* p) {
    if (fq_is_zero(&p->z) || fq_is_zero(&p->y)) { memset(r, 0, sizeof *r); return; }
    fq a, b, c, d, e, f, t;
    fq_sqr(&a, &p->x); fq_sqr(&b, &p->y); fq_sqr(&c, &b);
    fq_add(&t, &p->x, &b); fq_sqr(&t, &t); fq_sub(&t, &t, &a); fq_sub(&t, &t, &c); fq_add(&d, &t, &t);
    fq_add(&e, &a, &a); fq_add(&e, &e, &a); fq_sqr(&f, &e);
    fq z3; fq_mul(&z3, &p->y, &p->z); fq_add(&z3, &z3, &z3);
    fq x3; fq_sub(&x3, &f, &d); fq_sub(&x3, &x3, &d);
    fq c8; fq_add(&c8, &c, &c); fq_add(&c8, &c8, &c8); fq_add(&c8, &c8, &c8);
    fq y3; fq_sub(&t, &d, &x3); fq_mul(&y3, &e, &t); fq_sub(&y3, &y3, &c8);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add_affine(jac* r, const jac* p, const aff* q) {
    if (fq_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = ONE; return; }
    fq z1z1, u2, s2, h, rr, hh, hhh, v, t;
    fq_sqr(&z1z1, &p->z); fq_mul(&u2, &q->x, &z1z1); fq_mul(&s2, &q->y, &p->z); fq_mul(&s2, &s2, &z1z1);
    fq_sub(&h, &u2, &p->x); fq_sub(&rr, &s2, &p->y);
    if (fq_is_zero(&h)) { if (fq_is_zero(&rr)) { jac_dbl(r, p); } else { memset(r, 0, sizeof *r); } return; }
    fq_sqr(&hh, &h); fq_mul(&hhh, &h, &hh); fq_mul(&v, &p->x, &hh);
    fq x3; fq_sqr(&x3, &rr); fq_sub(&x3, &x3, &hhh); fq_sub(&x3, &x3, &v); fq_sub(&x3, &x3, &v);
    fq y3; fq_sub(&t, &v, &x3); fq_mul(&y3, &rr, &t); fq_mul(&t, &p->y, &hhh); fq_sub(&y3, &y3, &t);
    fq z3; fq_mul(&z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void jac_add(jac* r, const jac* p, const jac* q) {
    if (fq_is_zero(&p->z)) { *r = *q; return; }
    if (fq_is_zero(&q->z)) { *r = *p; return; }
    fq z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t;
    fq_sqr(&z1z1, &p->z); fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2); fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &q->z); fq_mul(&s1, &s1, &z2z2); fq_mul(&s2, &q->y, &p->z); fq_mul(&s2, &s2, &z1z1);
    fq_sub(&h, &u2, &u1); fq_sub(&rr, &s2, &s1);
    if (fq_is_zero(&h)) { if (fq_is_zero(&rr)) { jac_dbl(r, p); } else { memset(r, 0, sizeof *r); } return; }
    fq_sqr(&hh, &h); fq_mul(&hhh, &h, &hh); fq_mul(&v, &u1, &hh);
    fq x3; fq_sqr(&x3, &rr); fq_sub(&x3, &x3, &hhh); fq_sub(&x3, &x3, &v); fq_sub(&x3, &x3, &v);
    fq y3; fq_sub(&t, &v, &x3); fq_mul(&y3, &rr, &t); fq_mul(&t, &s1, &hhh); fq_sub(&y3, &y3, &t);
    fq z3; fq_mul(&z3, &p->z, &q->z); fq_mul(&z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

static unsigned window_digit(const uint8_t* k, unsigned bit, unsigned c) {
    unsigned d = 0;
    for (unsigned b = 0; b < c; b++) { unsigned pos = bit + b; if (pos < 256 && ((k[pos >> 3] >> (pos & 7)) & 1)) d |= 1u << b; }
    return d;
}

/* points: n x 96 bytes (x | y big-endian, affine, not infinity); scalars: n x 32 bytes little-endian (< 2^256).
 * out: 96-byte zcash uncompressed result (0x40 00.. for infinity).  Returns 0. */
int oracle_g1_msm(const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out) {
    unsigned c = 3;
    if (n >= 32) { c = 0; for (size_t t = n; t; t >>= 1) c++; c = c * 69 / 100 + 2; }
    if (c > 16) c = 16;
    unsigned nwin = (255 + c - 1) / c;
    aff* pts = (aff*)malloc(n * sizeof(aff));
    for (size_t i = 0; i < n; i++) { fq_from_be(&pts[i].x, points + 96 * i); fq_from_be(&pts[i].y, points + 96 * i + 48); }
    jac* buckets = (jac*)malloc(((size_t)1 << c) * sizeof(jac));
    jac total; memset(&total, 0, sizeof total);
    for (int w = (int)nwin - 1; w >= 0; w--) {
        for (unsigned k = 0; k < c; k++) jac_dbl(&total, &total);
        memset(buckets, 0, ((size_t)1 << c) * sizeof(jac));
        for (size_t i = 0; i < n; i++) {
            unsigned d = window_digit(scalars + 32 * i, (unsigned)w * c, c);
            if (d) jac_add_affine(&buckets[d], &buckets[d], &pts[i]);
        }
        jac run, acc; memset(&run, 0, sizeof run); memset(&acc, 0, sizeof acc);
        for (unsigned d = (1u << c) - 1; d >= 1; d--) { jac_add(&run, &run, &buckets[d]); jac_add(&acc, &acc, &run); }
        jac_add(&total, &total, &acc);
    }
    free(buckets); free(pts);
    if (fq_is_zero(&total.z)) { memset(out, 0, 96); out[0] = 0x40; return 0; }
    fq zi, zi2, zi3, x, y;
    fq_inv(&zi, &total.z); fq_sqr(&zi2, &zi); fq_mul(&zi3, &zi2, &zi);
    fq_mul(&x, &total.x, &zi2); fq_mul(&y, &total.y, &zi3);
    fq_to_be(out, &x); fq_to_be(out + 48, &y);
    (void)fq_eq;
    return 0;
}
