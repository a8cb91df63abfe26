// CRS coefficient oracle on the device (structs.rs:35-45,147-171).
//   coefficient(counter c) = sample(ChaCha20(key = big-endian bytes of (base_seed + c), block 0..))
// ChaCha20 as in rand_chacha 0.3.1 (20 rounds, 64-bit block counter in words 12-13, stream 0);
// sample = rand 0.8.5 UniformInt<i128>::sample_single for the range 0..Q: v = 128 keystream bits
// (4 little-endian u32 words, low first), (hi, lo) = v * Q as 256 bits, accept iff lo <= (Q << 115) - 1,
// i.e. iff the top 13 bits of lo are not all ones; return hi.  Rejected draws (probability 2^-13)
// continue with the next 128 keystream bits.
#pragma once
#include "lab_field.cuh"

struct LabSeed {
    uint64_t limb[4];   // the 256-bit big-endian base seed as an integer, limb[0] least significant
    uint32_t one;       // always 1, but only known at run time (kernel parameter / constant bank): a + b is written
                        // b * one + a so that ChaCha20's 320 additions per block issue as IMAD on the FMA pipe instead of
                        // IADD3 on the ALU pipe, which the xors and rotates already saturate
    uint32_t pad;
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lab_bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint32_t lab_rotl(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

// key words of ChaCha20Rng::from_seed(be_bytes(base + (chi:clo))) -- the counter's low 32 bits land
// byte-swapped in key word 7 (structs.rs:59-68,155-165; rand_chacha reads the seed as LE u32s)
__device__ __forceinline__ void lab_key_from_counter(const LabSeed &s, uint64_t clo, uint64_t chi, uint32_t (&key)[8]) {
    uint64_t s0 = s.limb[0] + clo;
    uint64_t c = s0 < clo;
    uint64_t s1 = s.limb[1] + chi;
    uint64_t c1 = s1 < chi;
    s1 += c;
    c1 += (s1 < c);
    uint64_t s2 = s.limb[2] + c1;
    uint64_t c2 = s2 < c1;
    uint64_t s3 = s.limb[3] + c2;
    key[0] = lab_bswap32((uint32_t)(s3 >> 32));
    key[1] = lab_bswap32((uint32_t)s3);
    key[2] = lab_bswap32((uint32_t)(s2 >> 32));
    key[3] = lab_bswap32((uint32_t)s2);
    key[4] = lab_bswap32((uint32_t)(s1 >> 32));
    key[5] = lab_bswap32((uint32_t)s1);
    key[6] = lab_bswap32((uint32_t)(s0 >> 32));
    key[7] = lab_bswap32((uint32_t)s0);
}

#define LAB_QR(a, b, c, d)                                                                       \
    a = b * one + a; d ^= a; d = lab_rotl(d, 16); c = d * one + c; b ^= c; b = lab_rotl(b, 12);  \
    a = b * one + a; d ^= a; d = lab_rotl(d, 8);  c = d * one + c; b ^= c; b = lab_rotl(b, 7);

// the 10 double rounds on NB independent states (interleaved for ILP); `one` == 1 (see LabSeed)
template <int NB>
__device__ __forceinline__ void lab_chacha_rounds(uint32_t (&x)[NB][16], const uint32_t one) {
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
            LAB_QR(x[b][0], x[b][4], x[b][8], x[b][12])
            LAB_QR(x[b][1], x[b][5], x[b][9], x[b][13])
            LAB_QR(x[b][2], x[b][6], x[b][10], x[b][14])
            LAB_QR(x[b][3], x[b][7], x[b][11], x[b][15])
        }
#pragma unroll
        for (int b = 0; b < NB; b++) {
            LAB_QR(x[b][0], x[b][5], x[b][10], x[b][15])
            LAB_QR(x[b][1], x[b][6], x[b][11], x[b][12])
            LAB_QR(x[b][2], x[b][7], x[b][8], x[b][13])
            LAB_QR(x[b][3], x[b][4], x[b][9], x[b][14])
        }
    }
}

__device__ __forceinline__ void lab_chacha_init(uint32_t (&x)[16], const uint32_t (&key)[8], uint64_t block) {
    x[0] = 0x61707865u; x[1] = 0x3320646eu; x[2] = 0x79622d32u; x[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) x[4 + i] = key[i];
    x[12] = (uint32_t)block; x[13] = (uint32_t)(block >> 32);
    x[14] = 0u; x[15] = 0u;
}

// rand 0.8.5 sample_single on four keystream words. Returns true when accepted.
__device__ __forceinline__ bool lab_sample_u128(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t &out) {
    uint64_t t = (uint64_t)w0 * LABQ;
    t = (uint64_t)w1 * LABQ + (t >> 32);
    t = (uint64_t)w2 * LABQ + (t >> 32);
    t = (uint64_t)w3 * LABQ + (t >> 32);
    out = (uint32_t)(t >> 32);
    return (((uint32_t)t) >> 19) != 0x1FFFu;
}

// slow path after `first_attempt` rejected draws: full blocks, any number of further attempts
__device__ __noinline__ uint32_t lab_crs_coeff_slow(const LabSeed &seed, uint64_t clo, uint64_t chi, uint32_t first_attempt) {
    uint32_t key[8];
    lab_key_from_counter(seed, clo, chi, key);
    for (uint32_t a = first_attempt;; a++) {
        uint32_t x[1][16], init[16];
        lab_chacha_init(x[0], key, (uint64_t)(a >> 2));
#pragma unroll
        for (int i = 0; i < 16; i++) init[i] = x[0][i];
        lab_chacha_rounds<1>(x, seed.one);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t v = x[0][i] + init[i];
            if ((i >> 2) == (int)(a & 3u)) w[i & 3] = v;
        }
        uint32_t out;
        if (lab_sample_u128(w[0], w[1], w[2], w[3], out)) return out;
    }
}

// NB coefficients at counters (chi:clo) + off[b]; only keystream words 0..3 of block 0 are finished
// on the fast path (the compiler drops the dead tail of the last round).
// Key setup: the 256-bit sum seed + (chi:clo) and the byte swaps of its upper six key words are done once for all NB
// blocks; a block only adds its small offset to the low 64-bit limb and swaps two words.  If that addition carries out
// of the low limb (possible only for seeds whose low limb is within `off` of 2^64) the block is recomputed by the
// generic slow path, which also handles rejected draws.
template <int NB>
__device__ __forceinline__ void lab_crs_coeffs(const LabSeed &seed, uint64_t clo, uint64_t chi, const uint32_t (&off)[NB], uint32_t (&out)[NB]) {
    const uint64_t s0 = seed.limb[0] + clo;
    const uint64_t c0 = s0 < clo;
    uint64_t s1 = seed.limb[1] + chi;
    uint64_t c1 = s1 < chi;
    s1 += c0;
    c1 += (s1 < c0);
    const uint64_t s2 = seed.limb[2] + c1;
    const uint64_t c2 = s2 < c1;
    const uint64_t s3 = seed.limb[3] + c2;
    uint32_t khi[6];
    khi[0] = lab_bswap32((uint32_t)(s3 >> 32));
    khi[1] = lab_bswap32((uint32_t)s3);
    khi[2] = lab_bswap32((uint32_t)(s2 >> 32));
    khi[3] = lab_bswap32((uint32_t)s2);
    khi[4] = lab_bswap32((uint32_t)(s1 >> 32));
    khi[5] = lab_bswap32((uint32_t)s1);
    uint32_t x[NB][16];
    bool carry[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const uint64_t t = s0 + off[b];
        carry[b] = t < s0;
        x[b][0] = 0x61707865u; x[b][1] = 0x3320646eu; x[b][2] = 0x79622d32u; x[b][3] = 0x6b206574u;
#pragma unroll
        for (int i = 0; i < 6; i++) x[b][4 + i] = khi[i];
        x[b][10] = lab_bswap32((uint32_t)(t >> 32));
        x[b][11] = lab_bswap32((uint32_t)t);
        x[b][12] = 0u; x[b][13] = 0u; x[b][14] = 0u; x[b][15] = 0u;
    }
    lab_chacha_rounds<NB>(x, seed.one);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        uint32_t w0 = x[b][0] + 0x61707865u, w1 = x[b][1] + 0x3320646eu, w2 = x[b][2] + 0x79622d32u, w3 = x[b][3] + 0x6b206574u;
        const bool ok = lab_sample_u128(w0, w1, w2, w3, out[b]);
        if (carry[b] || !ok) {
            uint64_t lo = clo + off[b];
            uint64_t hi = chi + (lo < clo);
            out[b] = lab_crs_coeff_slow(seed, lo, hi, carry[b] ? 0u : 1u);
        }
    }
}
#endif
