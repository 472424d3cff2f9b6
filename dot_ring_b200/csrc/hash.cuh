// SHA-512 and SHAKE128, host + device, one thread per message.
//
// The reference hashes with Python's hashlib: the VRF transcript (dot_ring/vrf/primitives.py:26-174,
// SHA-512 in counter mode), expand_message_xmd (dot_ring/curve/curve.py:145-185) and the ring-proof
// Fiat-Shamir transcript (dot_ring/ring_proof/transcript/transcript.py:21-136, SHAKE128 with
// non-finalising squeezes).  Moving them into the library keeps whole proof batches on the device
// between the arithmetic phases instead of bouncing through a per-proof Python loop.
#pragma once
#include <cstdint>

#include "fp.cuh"

namespace dr {

#define DR_SHA512_K_TABLE {                                                                                                          \
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL, 0x3956c25bf348b538ULL, \
    0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL, 0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, \
    0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL, 0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, \
    0xc19bf174cf692694ULL, 0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL, \
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL, 0x983e5152ee66dfabULL, \
    0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL, 0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, \
    0x06ca6351e003826fULL, 0x142929670a0e6e70ULL, 0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, \
    0x53380d139d95b3dfULL, 0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL, \
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL, 0xd192e819d6ef5218ULL, \
    0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL, 0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, \
    0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL, 0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, \
    0x682e6ff3d6b2b8a3ULL, 0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL, \
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL, 0xca273eceea26619cULL, \
    0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL, 0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, \
    0x113f9804bef90daeULL, 0x1b710b35131c471bULL, 0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, \
    0x431d67c49c100d4cULL, 0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL }
#define DR_KECCAK_RC_TABLE {                                                                                                          \
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, \
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, \
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, \
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL, \
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL }
static const uint64_t SHA512_K_HOST[80] = DR_SHA512_K_TABLE;
static const uint64_t KECCAK_RC_HOST[24] = DR_KECCAK_RC_TABLE;
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t SHA512_K_DEV[80] = DR_SHA512_K_TABLE;
static __device__ __constant__ uint64_t KECCAK_RC_DEV[24] = DR_KECCAK_RC_TABLE;
#endif
DR_HD uint64_t keccak_rc(int round) {
#if defined(__CUDA_ARCH__)
    return KECCAK_RC_DEV[round];
#else
    return KECCAK_RC_HOST[round];
#endif
}


// ---------------------------------------------------------------------------------------------
struct Sha512 {
    uint64_t h[8];
    uint8_t buf[128];
    uint32_t fill;
    uint64_t total;

    DR_HD static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

    DR_HD void init() {
        h[0] = 0x6a09e667f3bcc908ULL;
        h[1] = 0xbb67ae8584caa73bULL;
        h[2] = 0x3c6ef372fe94f82bULL;
        h[3] = 0xa54ff53a5f1d36f1ULL;
        h[4] = 0x510e527fade682d1ULL;
        h[5] = 0x9b05688c2b3e6c1fULL;
        h[6] = 0x1f83d9abfb41bd6bULL;
        h[7] = 0x5be0cd19137e2179ULL;
        fill = 0;
        total = 0;
    }

    // round constants: a table in constant memory (a function-local array indexed at run time is rebuilt on the thread's stack --
    // forty 128-bit local stores -- by every call; ncu showed those stores as the top stall of the one-thread-per-item kernels)
    DR_HD static uint64_t K(int i) {
#if defined(__CUDA_ARCH__)
        return SHA512_K_DEV[i];
#else
        return SHA512_K_HOST[i];
#endif
    }

    DR_HD_COLD void compress(const uint8_t* p) {
        uint64_t w[16];
        for (int i = 0; i < 16; i++) {
            uint64_t x = 0;
            for (int b = 0; b < 8; b++) x = (x << 8) | p[8 * i + b];
            w[i] = x;
        }
        uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        // 5 groups of 16 rounds, each group fully unrolled: the message-schedule window w[16] is indexed by compile-time constants
        // only and stays in registers
#pragma unroll 1
        for (int blk = 0; blk < 80; blk += 16) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                uint64_t wi;
                if (blk == 0) {
                    wi = w[k];
                } else {
                    uint64_t w15 = w[(k + 1) & 15], w2 = w[(k + 14) & 15];
                    uint64_t s0 = rotr(w15, 1) ^ rotr(w15, 8) ^ (w15 >> 7);
                    uint64_t s1 = rotr(w2, 19) ^ rotr(w2, 61) ^ (w2 >> 6);
                    wi = w[k] + s0 + w[(k + 9) & 15] + s1;
                    w[k] = wi;
                }
                uint64_t S1 = rotr(e, 14) ^ rotr(e, 18) ^ rotr(e, 41);
                uint64_t ch = (e & f) ^ (~e & g);
                uint64_t t1 = hh + S1 + ch + K(blk + k) + wi;
                uint64_t S0 = rotr(a, 28) ^ rotr(a, 34) ^ rotr(a, 39);
                uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
                uint64_t t2 = S0 + mj;
                hh = g;
                g = f;
                f = e;
                e = d + t1;
                d = c;
                c = b;
                b = a;
                a = t1 + t2;
            }
        }
        h[0] += a;
        h[1] += b;
        h[2] += c;
        h[3] += d;
        h[4] += e;
        h[5] += f;
        h[6] += g;
        h[7] += hh;
    }

    DR_HD void update(const uint8_t* data, uint32_t len) {
        total += len;
#pragma unroll 1
        for (uint32_t i = 0; i < len; i++) {
            buf[fill++] = data[i];
            if (fill == 128) {
                compress(buf);
                fill = 0;
            }
        }
    }
    DR_HD void update_byte(uint8_t b) { update(&b, 1); }

    DR_HD void final(uint8_t* out64) {
        uint64_t bits = total * 8;
        buf[fill++] = 0x80;
        if (fill > 112) {
            while (fill < 128) buf[fill++] = 0;
            compress(buf);
            fill = 0;
        }
        while (fill < 120) buf[fill++] = 0;  // upper 64 bits of the 128-bit length are zero
        for (int i = 0; i < 8; i++) buf[120 + i] = (uint8_t)(bits >> (56 - 8 * i));
        compress(buf);
        for (int i = 0; i < 8; i++)
            for (int b = 0; b < 8; b++) out64[8 * i + b] = (uint8_t)(h[i] >> (56 - 8 * b));
    }
};

// ---------------------------------------------------------------------------------------------
struct Shake128 {
    static constexpr int RATE = 168;
    uint64_t st[25];
    uint32_t pos;  // bytes absorbed into the current block

    DR_HD static uint64_t rotl(uint64_t x, int n) { return n ? ((x << n) | (x >> (64 - n))) : x; }

    DR_HD void init() {
        for (int i = 0; i < 25; i++) st[i] = 0;
        pos = 0;
    }

    DR_HD_COLD static void permute(uint64_t* a) {
        constexpr int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
        constexpr int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
        // the state is copied into locals and every inner loop is unrolled, so all indices are compile-time constants and the 25
        // lanes stay in registers (with rolled inner loops the state lived in local memory: ~4x the latency of a permutation,
        // which the one-thread-per-proof transcript kernels pay in full)
        uint64_t s[25];
#pragma unroll
        for (int i = 0; i < 25; i++) s[i] = a[i];
#pragma unroll 1
        for (int round = 0; round < 24; round++) {
            uint64_t bc[5];
#pragma unroll
            for (int i = 0; i < 5; i++) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
#pragma unroll
            for (int i = 0; i < 5; i++) {
                uint64_t t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1);
#pragma unroll
                for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
            }
            uint64_t t = s[1];
#pragma unroll
            for (int i = 0; i < 24; i++) {
                const int j = piln[i];
                uint64_t b0 = s[j];
                s[j] = rotl(t, rotc[i]);
                t = b0;
            }
#pragma unroll
            for (int j = 0; j < 25; j += 5) {
#pragma unroll
                for (int i = 0; i < 5; i++) bc[i] = s[j + i];
#pragma unroll
                for (int i = 0; i < 5; i++) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
            }
            s[0] ^= keccak_rc(round);
        }
#pragma unroll
        for (int i = 0; i < 25; i++) a[i] = s[i];
    }

    DR_HD void absorb(const uint8_t* data, uint32_t len) {
#pragma unroll 1
        for (uint32_t i = 0; i < len; i++) {
            st[pos >> 3] ^= (uint64_t)data[i] << (8 * (pos & 7));
            pos++;
            if (pos == RATE) {
                permute(st);
                pos = 0;
            }
        }
    }
    DR_HD void absorb_byte(uint8_t b) { absorb(&b, 1); }
    DR_HD void absorb_be32(uint32_t x) {
        uint8_t b[4] = {(uint8_t)(x >> 24), (uint8_t)(x >> 16), (uint8_t)(x >> 8), (uint8_t)x};
        absorb(b, 4);
    }

    // SHAKE128(everything absorbed so far)[0:outlen] without disturbing the absorbing state
    // (hashlib's digest() semantics, which the ring transcript relies on).
    DR_HD void squeeze_snapshot(uint8_t* out, uint32_t outlen) const {
        uint64_t s[25];
        for (int i = 0; i < 25; i++) s[i] = st[i];
        s[pos >> 3] ^= (uint64_t)0x1F << (8 * (pos & 7));
        s[(RATE - 1) >> 3] ^= (uint64_t)0x80 << (8 * ((RATE - 1) & 7));
        permute(s);
        uint32_t off = 0;
        for (uint32_t i = 0; i < outlen; i++) {
            if (off == RATE) {
                permute(s);
                off = 0;
            }
            out[i] = (uint8_t)(s[off >> 3] >> (8 * (off & 7)));
            off++;
        }
    }
};

}  // namespace dr
