// JL stage on a 2-bit packed projection matrix (proofgen.rs:429-457, util.rs:511-526, verification.rs:553-566).
//
// Packed format ("pi2"): the entries of one row of Pi_i are in {-1, 0, 1}; 16 consecutive entries c = 16 w + k share one
// 32-bit word:   bit k = (entry == +1),  bit 16 + k = (entry == -1).   A row of N*64 entries is N*4 words (2 bits per entry,
// a quarter of the int8 form), rows and witness vectors follow each other like in the int8 layout: pi2[R][256][N*4].
//
// k_jl2 -- p_j = sum_c Pi[j][c] s[c] by table lookup ("Four Russians"): a CTA takes one unit of 1024 coefficients of one
// witness vector, builds in shared memory, for each of its 128 groups of 8 coefficients, the 256 subset sums
// tab[g][m] = sum_{b in m} s[8 g + b] (16 bits: at most 8 * 8190), and then every row of Pi costs one lookup per byte of
// its masks: plus-byte lookups are added, minus-byte lookups subtracted.  Lane l owns groups 4 l .. 4 l + 3 and shared-memory
// bank l: entry (m, group q of the lane) sits at byte m * 256 + (q >> 1) * 128 + 4 l + 2 (q & 1), so neither the lookups
// (random m per lane) nor the table stores ever conflict, the address of an entry is one byte-permute of the mask word,
// and the table is written as 32-bit words holding two groups' sums, updated with one packed add per Gray-code step.
// The 256 rows are split over 8 warps x 32 rows; a lane keeps 32 row accumulators across all the units its CTA processes
// (persistent grid) and one butterfly of 31 shuffles at the very end turns them into per-row totals (lane l = row l),
// which go to the global int64 sums with one atomic per row and CTA.  Per entry the kernel does 1/8 lookup + 1/32 table
// store: it leans on the shared-memory pipe (ncu: LSU data pipe 77 % busy), at about half the rate HBM delivers the bytes.
//
// k_piT_omega2 -- v[c] = sum_j omega_j Pi[j][c] mod q (first half of phi'', proofgen.rs:244-253) from the same words: a
// thread owns one word (16 coefficients), walks the 256 rows (coalesced across threads) and accumulates omega split into
// two 7-bit limbs against x = (word >> k) & 0x00010001 = plus_k + 2^16 minus_k, so that one IMAD serves both masks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lab_field.cuh"

namespace lab {

constexpr int JL2_UNIT = 1024;                     // coefficients per table unit (128 groups of 8; 16-bit subset sums)
constexpr int JL2_THREADS = 256;                   // 8 warps x 32 rows = the 256 JL rows (verification.rs:559)
constexpr size_t JL2_SMEM = 2 * 256 * 32 * sizeof(uint32_t);   // 64 KB: three CTAs per SM

// int8 {-1,0,1} -> packed words.  n_words words of 16 entries each; thread per word, one 16-byte load.  Per four entries (one
// 32-bit word q of the input): lo = bit 0 of every byte (set for +1 = 0x01 and -1 = 0xFF), hi = bit 7 (set for -1 only), so
// plus = lo ^ hi and minus = hi; the four mask bits, 8 apart, are gathered into a nibble by one multiplication
// (x * 0x01020408 puts bit 8 i of x at bit 24 + i, no two partial products share a bit position).
__device__ __forceinline__ uint32_t jl2_nibble(uint32_t x) { return (x * 0x01020408u) >> 24 & 0xFu; }
__global__ void __launch_bounds__(256) k_pi_pack(const int8_t *__restrict__ pi, size_t n_words, uint32_t *__restrict__ out) {
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; w < n_words; w += stride) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(pi) + w);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
        uint32_t word = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint32_t lo = q[t] & 0x01010101u, hi = (q[t] >> 7) & 0x01010101u;
            word |= jl2_nibble(lo ^ hi) << (4 * t) | jl2_nibble(hi) << (16 + 4 * t);
        }
        out[w] = word;
    }
}
// packed words -> int8 (the transcript's pi_i_all needs the entries back: proofgen.rs:445-453)
__global__ void __launch_bounds__(256) k_pi_unpack(const uint32_t *__restrict__ pi2, size_t n_words, int8_t *__restrict__ out) {
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; w < n_words; w += stride) {
        const uint32_t word = pi2[w];
        uint32_t q[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            q[t] = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int k = 4 * t + e;
                const uint32_t b = (word >> k & 1u) ? 1u : ((word >> (16 + k) & 1u) ? 0xFFu : 0u);
                q[t] |= b << (8 * e);
            }
        }
        reinterpret_cast<uint4 *>(out)[w] = make_uint4(q[0], q[1], q[2], q[3]);
    }
}

// Shared-memory byte address of table entry (group q of the lane, subset m): m * 256 + (q >> 1) * 128 + lane * 4 + (q & 1) * 2, formed
// by ONE byte permute: result byte 0 = byte 0 of `off` (the lane's constant part, < 256), byte 1 = byte BYTE of the mask word.
template <int BYTE>
__device__ __forceinline__ uint32_t jl2_addr(uint32_t x, uint32_t off) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(off), "n"(0x7704 | (BYTE << 4)));   // selectors: [3]=off.b3 (0) [2]=off.b3 (0) [1]=x.bBYTE [0]=off.b0
    return r;
}
__device__ __forceinline__ int jl2_lds16(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return (int)v;
}

// pi2: rows of this call's witness vectors, [ni][256][W] with W = ND / 16 words per row; S: the full witness [R][ND] (device);
// vector li of pi2 is witness vector i0 + li.  p: int64[256], accumulated with atomics (zero it first).
__global__ void __launch_bounds__(JL2_THREADS, 3) k_jl2(const uint32_t *__restrict__ pi2, const uint32_t *__restrict__ S, uint64_t ND, uint32_t W,
                                                         uint32_t i0, uint32_t units_per_vec, uint64_t total_units, unsigned long long *__restrict__ p) {
    extern __shared__ __align__(256) uint32_t tab[];            // [m][group pair][32 lanes] words of two 16-bit subset sums: 256 bytes per subset m
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t off0 = (uint32_t)lane * 4u, off1 = off0 + 2u, off2 = off0 + 128u, off3 = off0 + 130u;
    int acc[32];
#pragma unroll
    for (int r = 0; r < 32; r++) acc[r] = 0;
    for (uint64_t unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const uint32_t li = (uint32_t)(unit / units_per_vec), u = (uint32_t)(unit % units_per_vec);
        // ---- the lane's 32 coefficients = groups 0..3 of the lane, canonical, as packed pairs: svp[p][k] = s[16 p + k] | s[16 p + 8 + k] << 16
        //      (bit k of group 2 p in the low half, of group 2 p + 1 in the high half); past the end of the vector: zero ----
        uint32_t svp[2][8];
        {
            const uint64_t c0 = (uint64_t)u * JL2_UNIT + (uint64_t)lane * 32;
            const bool in = c0 < ND;                                    // ND is a multiple of 64: all 32 or none
            const uint4 *src = reinterpret_cast<const uint4 *>(S + (uint64_t)(i0 + li) * ND + (in ? c0 : 0));
#pragma unroll
            for (int pp = 0; pp < 2; pp++) {
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; q++) v[q] = in ? __ldg(src + 4 * pp + q) : make_uint4(0, 0, 0, 0);
                svp[pp][0] = lab_canon(v[0].x) | lab_canon(v[2].x) << 16; svp[pp][1] = lab_canon(v[0].y) | lab_canon(v[2].y) << 16;
                svp[pp][2] = lab_canon(v[0].z) | lab_canon(v[2].z) << 16; svp[pp][3] = lab_canon(v[0].w) | lab_canon(v[2].w) << 16;
                svp[pp][4] = lab_canon(v[1].x) | lab_canon(v[3].x) << 16; svp[pp][5] = lab_canon(v[1].y) | lab_canon(v[3].y) << 16;
                svp[pp][6] = lab_canon(v[1].z) | lab_canon(v[3].z) << 16; svp[pp][7] = lab_canon(v[1].w) | lab_canon(v[3].w) << 16;
            }
        }
        // The unit's two words (32 entries) of rows 32 w .. 32 w + 7 are requested before the table is built.  A lane outside the
        // row (ragged last unit) reads the first words of the unit instead and masks them: no predicated address arithmetic.
        const bool win = (uint64_t)u * (JL2_UNIT / 16) + 2 * lane < W;
        const uint2 *row = reinterpret_cast<const uint2 *>(pi2 + ((uint64_t)li * 256 + (uint64_t)w * 32) * W + (uint64_t)u * (JL2_UNIT / 16) + (win ? 2 * lane : 0));
        const uint32_t keep = win ? 0xFFFFFFFFu : 0u, W2 = W / 2;
        uint2 wd[8];
#pragma unroll
        for (int r = 0; r < 8; r++) wd[r] = __ldg(row + (uint32_t)r * W2);
        __syncthreads();                                                // every warp is done with the previous table
        // ---- subset sums: warp w fills m = 32 w + k (k in Gray-code order: one packed add or subtract per pair of entries; a field
        //      never leaves [0, 8 * 8190], so the two 16-bit halves never borrow from each other) ----
#pragma unroll
        for (int pp = 0; pp < 2; pp++) {
            const uint32_t *s8 = svp[pp];
            uint32_t val = ((w & 1) ? s8[5] : 0u) + ((w & 2) ? s8[6] : 0u) + ((w & 4) ? s8[7] : 0u);
            uint32_t *t = tab + (size_t)(w * 32) * 64 + pp * 32 + lane;
            t[0] = val;
            int idx = 0;
#pragma unroll
            for (int k = 1; k < 32; k++) {
                const int bit = (k & 1) ? 0 : ((k & 2) ? 1 : ((k & 4) ? 2 : ((k & 8) ? 3 : 4)));
                idx ^= 1 << bit;
                if (idx >> bit & 1) val += s8[bit];
                else val -= s8[bit];
                t[idx * 64] = val;
            }
        }
        __syncthreads();
        // ---- lookups: rows 32 w .. 32 w + 31 in four quarters of 8 rows; the next quarter's words are in flight meanwhile ----
#pragma unroll
        for (int qr = 0; qr < 4; qr++) {
            uint2 nx[8];
            if (qr < 3) {
#pragma unroll
                for (int r = 0; r < 8; r++) nx[r] = __ldg(row + (uint32_t)(8 * (qr + 1) + r) * W2);
            }
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const uint32_t x = wd[r].x & keep, y = wd[r].y & keep;
                const int a0 = jl2_lds16(tab_s + jl2_addr<0>(x, off0));   // plus masks: groups 0, 1 (word x), 2, 3 (word y)
                const int a1 = jl2_lds16(tab_s + jl2_addr<1>(x, off1));
                const int a2 = jl2_lds16(tab_s + jl2_addr<0>(y, off2));
                const int a3 = jl2_lds16(tab_s + jl2_addr<1>(y, off3));
                const int m0 = jl2_lds16(tab_s + jl2_addr<2>(x, off0));   // minus masks
                const int m1 = jl2_lds16(tab_s + jl2_addr<3>(x, off1));
                const int m2 = jl2_lds16(tab_s + jl2_addr<2>(y, off2));
                const int m3 = jl2_lds16(tab_s + jl2_addr<3>(y, off3));
                acc[8 * qr + r] += ((a0 - m0) + (a1 - m1)) + ((a2 - m2) + (a3 - m3));
            }
            if (qr < 3) {
#pragma unroll
                for (int r = 0; r < 8; r++) wd[r] = nx[r];
            }
        }
    }
    // ---- 32 accumulators per lane -> one row total per lane (row 32 w + lane) ----
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; k++) {
            const int send = up ? acc[k] : acc[k + s];
            const int keep2 = up ? acc[k + s] : acc[k];
            acc[k] = keep2 + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    if (acc[0]) atomicAdd(p + w * 32 + lane, (unsigned long long)(long long)acc[0]);
}

// v[i][c] = sum_j omega_j Pi_i[j][c] mod q from packed words.  total_words = R * W.
__global__ void __launch_bounds__(256) k_piT_omega2(const uint32_t *__restrict__ pi2, const uint32_t *__restrict__ omega, uint64_t total_words, uint32_t W,
                                                    uint32_t *__restrict__ v) {
    __shared__ uint32_t som[2][256];                 // omega_j = lo + 128 hi
    {
        const uint32_t o = lab_canon(omega[threadIdx.x]);
        som[0][threadIdx.x] = o & 127u;
        som[1][threadIdx.x] = o >> 7;
    }
    __syncthreads();
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_words) return;
    const uint64_t i = idx / W, wq = idx % W;
    const uint32_t *col = pi2 + i * 256 * (uint64_t)W + wq;
    uint32_t lo[16], hi[16];                         // packed: low half = sum over plus entries, high half = sum over minus entries (< 2^15 each)
#pragma unroll
    for (int k = 0; k < 16; k++) { lo[k] = 0; hi[k] = 0; }
#pragma unroll 4
    for (int j = 0; j < 256; j++) {
        const uint32_t x = __ldg(col + (uint64_t)j * W);
        const uint32_t ol = som[0][j], oh = som[1][j];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t t = (x >> k) & 0x00010001u;
            lo[k] += ol * t;
            hi[k] += oh * t;
        }
    }
    uint32_t out[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // (lo+ - lo-) + 128 (hi+ - hi-), each part < 2^15: keep everything positive before reducing
        const uint32_t pos = (lo[k] & 0xFFFFu) + 128u * (hi[k] & 0xFFFFu);
        const uint32_t neg = (lo[k] >> 16) + 128u * (hi[k] >> 16);
        out[k] = lab_canon(pos + 257u * LABQ - neg);          // neg <= 256 * 127 + 128 * 256 * 63 < 257 q
    }
    uint4 *dst = reinterpret_cast<uint4 *>(v + idx * 16);
#pragma unroll
    for (int q = 0; q < 4; q++) dst[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

// JL matrix entries straight into the packed form, same PRG stream and values as k_synth_pi (two bits per entry from PRG
// word e / 32: 0 -> -1, 3 -> +1, 1 and 2 -> 0; verification.rs:553-566).  Thread per PRG word = two packed words.
__device__ __forceinline__ uint64_t jl2_prg_u64(uint64_t base, uint64_t idx) {
    uint64_t z = base + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) k_synth_pi2(uint64_t base, size_t n_prg_words, uint32_t *__restrict__ out) {
    size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; wd < n_prg_words; wd += stride) {
        const uint64_t bits = jl2_prg_u64(base, wd);
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            const unsigned a = (unsigned)(bits >> (2 * t)) & 3u, b = (unsigned)(bits >> (2 * (t + 16))) & 3u;
            w0 |= (a == 3u ? 1u : 0u) << t | (a == 0u ? 1u : 0u) << (16 + t);
            w1 |= (b == 3u ? 1u : 0u) << t | (b == 0u ? 1u : 0u) << (16 + t);
        }
        reinterpret_cast<uint2 *>(out)[wd] = make_uint2(w0, w1);
    }
}

// ---- helpers of the in-library collectives (int64 sums over ranks, then mod q) ----
__global__ void k_widen_u32_i64(const uint32_t *__restrict__ in, size_t n, long long *__restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n; idx += stride) out[idx] = (long long)in[idx];
}
__global__ void k_modq_i64_u32(const long long *__restrict__ in, size_t n, uint32_t *__restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n; idx += stride) {
        long long m = in[idx] % (long long)LABQ;
        out[idx] = (uint32_t)(m < 0 ? m + (long long)LABQ : m);
    }
}

}  // namespace lab
