// CRS coefficient oracle on the device (structs.rs:35-45,147-171).
//   coefficient(counter c) = sample(ChaCha20(key = big-endian bytes of (base_seed + c), block 0..))
// ChaCha20 as in rand_chacha 0.3.1 (20 rounds, 64-bit block counter in words 12-13, stream 0);
// sample = rand 0.8.5 UniformInt<i128>::sample_single for the range 0..Q: v = 128 keystream bits
// (4 little-endian u32 words, low first), (hi, lo) = v * Q as 256 bits, accept iff lo <= (Q << 115) - 1,
// i.e. iff the top 13 bits of lo are not all ones; return hi.  Rejected draws (probability 2^-13)
// continue with the next 128 keystream bits.  (lab_sample_u128 is this rule word for word; lab_sample_w3 is the
// shortcut through the top word that the fast paths use, falling back to the former when word 3 does not decide.)
//
// One ChaCha20 block per 13-bit coefficient makes every CRS-regenerating kernel INT32-bound, so the block function is
// trimmed to what the result depends on:
//   * only key word 7 (the byte-swapped low 32 bits of seed + counter) differs between neighbouring coefficients, so the
//     part of the first double round that does not depend on it is computed once per 2^32 counters (LabHoist);
//   * the sample is decided by keystream word 3 alone except with probability 2^-13 (lab_sample_w3), so only the cone of
//     x3 is computed: the last diagonal round is one quarter round cut at its second `a` update, and the last column round
//     stops at the four words that quarter round reads (b of column 0, c of column 1, d of column 2, a of column 3);
//   * xor and rotate exist only on the ALU pipe; the additions are written b * one + a (IMAD, FMA pipe) and a
//     compile-time mask moves selected rotations to the FMA pipe as rotl(x, n) = hi32(x * 2^n) + lo32(x * 2^n)
//     (IMAD + IMAD.HI) until both pipes carry the same load.
#pragma once
#include "lab_field.cuh"

struct LabSeed {
    uint64_t limb[4];   // the 256-bit big-endian base seed as an integer, limb[0] least significant
    uint32_t one;       // always 1, but only known at run time (kernel parameter / constant bank): a + b is written
                        // b * one + a so that ChaCha20's additions issue as IMAD on the FMA pipe instead of
                        // IADD3 on the ALU pipe, which the xors and rotates already saturate
    uint32_t p16, p12, p8, p7;   // 2^16, 2^12, 2^8, 2^7, run-time for the same reason (FMA-pipe rotations)
    uint32_t pad[3];
};

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lab_bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ uint32_t lab_rotl(uint32_t x, int n) { return __funnelshift_l(x, x, n); }
// x + y on the FMA pipe
#ifdef LAB_NATIVE_ADD
__device__ __forceinline__ uint32_t lab_addf(uint32_t x, uint32_t y, uint32_t) { return x + y; }
#else
__device__ __forceinline__ uint32_t lab_addf(uint32_t x, uint32_t y, uint32_t one) { return y * one + x; }
#endif
// rotl on the FMA pipe: pw == 2^n at run time
__device__ __forceinline__ uint32_t lab_rotl_fma(uint32_t x, uint32_t pw) {
    uint32_t lo = x * pw, r;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(pw), "r"(lo));
    return r;
}
template <int N, bool FMA>
__device__ __forceinline__ uint32_t lab_rot(uint32_t x, const LabSeed &s) {
    if (FMA) return lab_rotl_fma(x, N == 16 ? s.p16 : (N == 12 ? s.p12 : (N == 8 ? s.p8 : s.p7)));
    return lab_rotl(x, N);
}

// quarter round; M = 4-bit mask, bit j set -> rotation j (16, 12, 8, 7) runs on the FMA pipe
template <uint32_t M>
__device__ __forceinline__ void lab_qr(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, const LabSeed &s) {
    a = lab_addf(a, b, s.one); d = lab_rot<16, (M & 1u) != 0>(d ^ a, s);
    c = lab_addf(c, d, s.one); b = lab_rot<12, (M & 2u) != 0>(b ^ c, s);
    a = lab_addf(a, b, s.one); d = lab_rot<8, (M & 4u) != 0>(d ^ a, s);
    c = lab_addf(c, d, s.one); b = lab_rot<7, (M & 8u) != 0>(b ^ c, s);
}
// the same, stopping once `a` is final (last diagonal round: only x0..x3 are consumed)
template <uint32_t M>
__device__ __forceinline__ void lab_qr_a_only(uint32_t &a, uint32_t b, uint32_t c, uint32_t d, const LabSeed &s) {
    a = lab_addf(a, b, s.one); d = lab_rot<16, (M & 1u) != 0>(d ^ a, s);
    c = lab_addf(c, d, s.one); b = lab_rot<12, (M & 2u) != 0>(b ^ c, s);
    a = lab_addf(a, b, s.one);
}

// the same, stopping once `c` is final (skips the last update of b) / once `d` is final (skips the last c and b)
template <uint32_t M>
__device__ __forceinline__ void lab_qr_c_only(uint32_t a, uint32_t b, uint32_t &c, uint32_t d, const LabSeed &s) {
    a = lab_addf(a, b, s.one); d = lab_rot<16, (M & 1u) != 0>(d ^ a, s);
    c = lab_addf(c, d, s.one); b = lab_rot<12, (M & 2u) != 0>(b ^ c, s);
    a = lab_addf(a, b, s.one); d = lab_rot<8, (M & 4u) != 0>(d ^ a, s);
    c = lab_addf(c, d, s.one);
}
template <uint32_t M>
__device__ __forceinline__ void lab_qr_d_only(uint32_t a, uint32_t b, uint32_t c, uint32_t &d, const LabSeed &s) {
    a = lab_addf(a, b, s.one); d = lab_rot<16, (M & 1u) != 0>(d ^ a, s);
    c = lab_addf(c, d, s.one); b = lab_rot<12, (M & 2u) != 0>(b ^ c, s);
    a = lab_addf(a, b, s.one); d = lab_rot<8, (M & 4u) != 0>(d ^ a, s);
}

// one double round on NB interleaved states; RM = 32-bit mask, nibble q = lab_qr mask of quarter round q
// (0..3 column, 4..7 diagonal)
template <int NB, uint32_t RM>
__device__ __forceinline__ void lab_double_round(uint32_t (&x)[NB][16], const LabSeed &s) {
#pragma unroll
    for (int b = 0; b < NB; b++) {
        lab_qr<(RM >> 0) & 15u>(x[b][0], x[b][4], x[b][8], x[b][12], s);
        lab_qr<(RM >> 4) & 15u>(x[b][1], x[b][5], x[b][9], x[b][13], s);
        lab_qr<(RM >> 8) & 15u>(x[b][2], x[b][6], x[b][10], x[b][14], s);
        lab_qr<(RM >> 12) & 15u>(x[b][3], x[b][7], x[b][11], x[b][15], s);
    }
#pragma unroll
    for (int b = 0; b < NB; b++) {
        lab_qr<(RM >> 16) & 15u>(x[b][0], x[b][5], x[b][10], x[b][15], s);
        lab_qr<(RM >> 20) & 15u>(x[b][1], x[b][6], x[b][11], x[b][12], s);
        lab_qr<(RM >> 24) & 15u>(x[b][2], x[b][7], x[b][8], x[b][13], s);
        lab_qr<(RM >> 28) & 15u>(x[b][3], x[b][4], x[b][9], x[b][14], s);
    }
}

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

constexpr uint32_t LAB_CC0 = 0x61707865u, LAB_CC1 = 0x3320646eu, LAB_CC2 = 0x79622d32u, LAB_CC3 = 0x6b206574u;

// The part of the first double round that is the same for all counters sharing everything but the low 32 bits of
// seed + counter (i.e. key words 0..6).  State after the first column round: columns 0..2 complete (A*), column 3 =
// QR(c3, k3, k7, 0) starts with P0 = c3 + k3 and P1 = rotl(P0, 16); the diagonal round then starts from
// Q0 = A0 + A5, Q1 = A1 + A6, Q2 = rotl(A12 ^ Q1, 16).
struct LabHoist {
    uint64_t tag_lo, tag_hi;   // (carry out of the low limb : high 32 bits of the low limb), counter high limb
    uint32_t k3, P0, P1;
    uint32_t Q0, A5, A10;
    uint32_t Q1, Q2, A6;
    uint32_t A2, A8, A13;
    uint32_t A4, A9, A14;
};
__device__ __forceinline__ void lab_hoist_invalidate(LabHoist &h) { h.tag_lo = ~0ull; h.tag_hi = ~0ull; }

// recompute for the 256-bit sum seed + (chi:clo): runs once per 2^32 counters
__device__ __forceinline__ void lab_hoist_compute(const LabSeed &seed, uint64_t clo, uint64_t chi, LabHoist &h) {
    uint32_t key[8];
    lab_key_from_counter(seed, clo, chi, key);
    const uint64_t s0 = seed.limb[0] + clo;
    h.tag_lo = ((uint64_t)(s0 < clo) << 32) | (s0 >> 32);
    h.tag_hi = chi;
    uint32_t a0 = LAB_CC0, a4 = key[0], a8 = key[4], a12 = 0u;
    uint32_t a1 = LAB_CC1, a5 = key[1], a9 = key[5], a13 = 0u;
    uint32_t a2 = LAB_CC2, a6 = key[2], a10 = key[6], a14 = 0u;
    lab_qr<0>(a0, a4, a8, a12, seed);
    lab_qr<0>(a1, a5, a9, a13, seed);
    lab_qr<0>(a2, a6, a10, a14, seed);
    h.k3 = key[3];
    h.P0 = LAB_CC3 + key[3];
    h.P1 = lab_rotl(h.P0, 16);
    h.Q0 = a0 + a5; h.A5 = a5; h.A10 = a10;
    h.Q1 = a1 + a6; h.Q2 = lab_rotl(a12 ^ h.Q1, 16); h.A6 = a6;
    h.A2 = a2; h.A8 = a8; h.A13 = a13;
    h.A4 = a4; h.A9 = a9; h.A14 = a14;
}
// make h valid for counters (chi:clo) + [0, 2^32 - low32(seed + clo))
__device__ __forceinline__ void lab_hoist_update(const LabSeed &seed, uint64_t clo, uint64_t chi, LabHoist &h) {
    const uint64_t s0 = seed.limb[0] + clo;
    const uint64_t tag = ((uint64_t)(s0 < clo) << 32) | (s0 >> 32);
    if (tag != h.tag_lo || chi != h.tag_hi) lab_hoist_compute(seed, clo, chi, h);
}

// keystream word 3 of block 0 (the top 32 bits of the 128-bit draw) for NB keys that share h and differ in key word 7
// UNR: unroll factor of the loop over double rounds 2..9 (1, 2, 4 or 8)
template <int NB, uint32_t RM, int UNR = 1>
__device__ __forceinline__ void lab_chacha_w3(const LabSeed &s, const LabHoist &h, const uint32_t (&k7)[NB], uint32_t (&w3)[NB]) {
    uint32_t x[NB][16];
    const uint32_t one = s.one;
    // ---- first double round, hoisted ----
#pragma unroll
    for (int b = 0; b < NB; b++) {
        // column 3: QR(c3, k3, k7, 0) from its third operation on
        uint32_t c = lab_addf(k7[b], h.P1, one);
        uint32_t bb = lab_rot<12, ((RM >> 12) & 2u) != 0>(h.k3 ^ c, s);
        uint32_t a = lab_addf(h.P0, bb, one);
        uint32_t d = lab_rot<8, ((RM >> 12) & 4u) != 0>(h.P1 ^ a, s);
        c = lab_addf(c, d, one);
        bb = lab_rot<7, ((RM >> 12) & 8u) != 0>(bb ^ c, s);
        x[b][3] = a; x[b][7] = bb; x[b][11] = c; x[b][15] = d;
    }
#pragma unroll
    for (int b = 0; b < NB; b++) {
        {   // diagonal 0: (A0, A5, A10, x15), a + b = Q0 known
            uint32_t d = lab_rot<16, ((RM >> 16) & 1u) != 0>(x[b][15] ^ h.Q0, s);
            uint32_t c = lab_addf(h.A10, d, one);
            uint32_t bb = lab_rot<12, ((RM >> 16) & 2u) != 0>(h.A5 ^ c, s);
            uint32_t a = lab_addf(h.Q0, bb, one);
            d = lab_rot<8, ((RM >> 16) & 4u) != 0>(d ^ a, s);
            c = lab_addf(c, d, one);
            bb = lab_rot<7, ((RM >> 16) & 8u) != 0>(bb ^ c, s);
            x[b][0] = a; x[b][5] = bb; x[b][10] = c; x[b][15] = d;
        }
        {   // diagonal 1: (A1, A6, x11, A12), a = Q1 and d = Q2 known
            uint32_t c = lab_addf(x[b][11], h.Q2, one);
            uint32_t bb = lab_rot<12, ((RM >> 20) & 2u) != 0>(h.A6 ^ c, s);
            uint32_t a = lab_addf(h.Q1, bb, one);
            uint32_t d = lab_rot<8, ((RM >> 20) & 4u) != 0>(h.Q2 ^ a, s);
            c = lab_addf(c, d, one);
            bb = lab_rot<7, ((RM >> 20) & 8u) != 0>(bb ^ c, s);
            x[b][1] = a; x[b][6] = bb; x[b][11] = c; x[b][12] = d;
        }
        x[b][2] = h.A2; x[b][8] = h.A8; x[b][13] = h.A13;
        lab_qr<(RM >> 24) & 15u>(x[b][2], x[b][7], x[b][8], x[b][13], s);
        x[b][4] = h.A4; x[b][9] = h.A9; x[b][14] = h.A14;
        lab_qr<(RM >> 28) & 15u>(x[b][3], x[b][4], x[b][9], x[b][14], s);
    }
    // ---- double rounds 2..9 ----
#pragma unroll UNR
    for (int r = 0; r < 8; r++) lab_double_round<NB, RM>(x, s);
    // ---- double round 10, only the cone of x3: the diagonal quarter round (x3, x4, x9, x14) up to its second `a` update, and
    //      of the column round what that reads -- column 0 up to b (all of it), column 1 up to c, column 2 up to d, column 3 up to a.
    //      576 xor / rotate operations per block depend on key word 7 in this form (640 in the plain block).
#pragma unroll
    for (int b = 0; b < NB; b++) {
        lab_qr<(RM >> 0) & 15u>(x[b][0], x[b][4], x[b][8], x[b][12], s);
        lab_qr_c_only<(RM >> 4) & 15u>(x[b][1], x[b][5], x[b][9], x[b][13], s);
        lab_qr_d_only<(RM >> 8) & 15u>(x[b][2], x[b][6], x[b][10], x[b][14], s);
        lab_qr_a_only<(RM >> 12) & 15u>(x[b][3], x[b][7], x[b][11], x[b][15], s);
    }
#pragma unroll
    for (int b = 0; b < NB; b++) {
        lab_qr_a_only<(RM >> 28) & 15u>(x[b][3], x[b][4], x[b][9], x[b][14], s);
        w3[b] = lab_addf(x[b][3], LAB_CC3, one);
    }
}

// the 10 double rounds on one full state (generic path: any block index, all 16 words)
__device__ __forceinline__ void lab_chacha_rounds1(uint32_t (&x)[1][16], const LabSeed &s) {
#pragma unroll 1
    for (int r = 0; r < 10; r++) lab_double_round<1, 0u>(x, s);
}

__device__ __forceinline__ void lab_chacha_init(uint32_t (&x)[16], const uint32_t (&key)[8], uint64_t block) {
    x[0] = LAB_CC0; x[1] = LAB_CC1; x[2] = LAB_CC2; x[3] = LAB_CC3;
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
    return ((uint32_t)t) < 0xFFF80000u;      // top 13 bits of the low half not all ones
}

// The same decision from the TOP keystream word alone.  v * Q = (w3 * Q) * 2^96 + rest * Q with rest = w2:w1:w0 < 2^96, so
// floor(rest * Q / 2^96) <= Q - 1: with L = lo32(w3 * Q) the top 32 bits of the low half of the product lie in
// [L, L + Q - 1].  When that interval ends below the rejection zone 0xFFF80000 (and therefore below 2^32: no carry into the
// high half) the draw is accepted with the value hi32(w3 * Q) whatever words 0..2 are.  Otherwise -- probability
// 2^-13 (1 + 2^-6) -- the caller recomputes the coefficient on the generic path from draw 0.
__device__ __forceinline__ bool lab_sample_w3(uint32_t w3, uint32_t &out) {
    const uint64_t t = (uint64_t)w3 * LABQ;
    out = (uint32_t)(t >> 32);
    return (uint32_t)t < 0xFFF80000u - (LABQ - 1u);
}

// generic path: any counter, any number of rejected draws (`first_attempt` of them already known to be rejected).
// Also the independent cross-check of the trimmed fast path: tests compare both on the same counters.
__device__ __forceinline__ uint32_t lab_crs_coeff_generic(const LabSeed &seed, uint64_t clo, uint64_t chi, uint32_t first_attempt) {
    uint32_t key[8];
    lab_key_from_counter(seed, clo, chi, key);
    for (uint32_t a = first_attempt;; a++) {
        uint32_t x[1][16], init[16];
        lab_chacha_init(x[0], key, (uint64_t)(a >> 2));
#pragma unroll
        for (int i = 0; i < 16; i++) init[i] = x[0][i];
        lab_chacha_rounds1(x, seed);
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

__device__ __noinline__ uint32_t lab_crs_coeff_slow(const LabSeed &seed, uint64_t clo, uint64_t chi, uint32_t first_attempt) {
    return lab_crs_coeff_generic(seed, clo, chi, first_attempt);
}

// NB coefficients at counters (chi:clo) + off[b].  h is the caller's hoist cache (lab_hoist_invalidate once, then reuse
// across calls: it is refreshed here when the high part of seed + counter changes).  A block whose offset carries out
// of the low 32 bits of seed + clo, or whose first draw word 3 alone does not decide, is recomputed by the generic path.
template <int NB, uint32_t RM, int UNR = 1>
__device__ __forceinline__ void lab_crs_coeffs(const LabSeed &seed, LabHoist &h, uint64_t clo, uint64_t chi, const uint32_t (&off)[NB], uint32_t (&out)[NB]) {
    lab_hoist_update(seed, clo, chi, h);
    const uint32_t lo32 = (uint32_t)seed.limb[0] + (uint32_t)clo;
    uint32_t k7[NB], w3[NB];
    bool slow[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const uint32_t t = lo32 + off[b];
        slow[b] = t < off[b];
        k7[b] = lab_bswap32(t);
    }
    lab_chacha_w3<NB, RM, UNR>(seed, h, k7, w3);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const bool ok = lab_sample_w3(w3[b], out[b]);
        if (slow[b] || !ok) {
            uint64_t lo = clo + off[b];
            uint64_t hi = chi + (lo < clo);
            out[b] = lab_crs_coeff_slow(seed, lo, hi, 0u);
        }
    }
}
#endif
