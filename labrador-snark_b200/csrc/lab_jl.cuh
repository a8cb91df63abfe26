// JL stage on a 2-bit packed projection matrix (proofgen.rs:429-457, util.rs:511-526, verification.rs:553-566).
//
// Packed format ("pi2"): the entries of one row of Pi_i are in {-1, 0, 1}; 16 consecutive entries c = 16 w + k share one
// 32-bit word:   bit k = (entry == +1),  bit 16 + k = (entry == -1).   A row of N*64 entries is N*4 words (2 bits per entry,
// a quarter of the int8 form), rows and witness vectors follow each other like in the int8 layout: pi2[R][256][N*4].
//
// k_jl2 -- p_j = sum_c Pi[j][c] s[c] by table lookup ("Four Russians"): a CTA takes one unit of 512 coefficients of one
// witness vector, builds in shared memory, for each of its 64 groups of 8 coefficients, the 256 subset sums
// tab[g][m] = sum_{b in m} s[8 g + b], and then every row of Pi costs one lookup per byte of its masks: plus-byte lookups
// are added, minus-byte lookups subtracted.  Layout tab[m][g & 1][g >> 1]: lane l owns groups 2 l, 2 l + 1 and
// shared-memory bank l, so neither the lookups (random m per lane) nor the table stores ever conflict, and the byte address
// m * 256 + (g & 1) * 128 + 4 l of an entry is one byte-permute of the mask word.  The 256 rows are
// split over 8 warps x 32 rows; a lane keeps 32 row accumulators across all the units its CTA processes (persistent grid)
// and one butterfly of 31 shuffles at the very end turns them into per-row totals (lane l = row l), which go to the global
// int64 sums with one atomic per row and CTA.  Per entry the kernel does 1/8 lookup + 1/16 table store: the shared-memory
// pipe, not the ALU pipe, is what it leans on (ncu: LSU data pipe 77 % busy), at half the rate HBM delivers the packed bytes.
// (A variant with 16-bit sums, 1024-coefficient units and 8-byte row loads -- 25 % fewer shared-memory wavefronts on paper --
// measured 3 % slower on a B200 and was dropped.)
//
// k_piT_omega2 -- v[c] = sum_j omega_j Pi[j][c] mod q (first half of phi'', proofgen.rs:244-253) from the same words: a
// thread owns one word (16 coefficients), walks the 256 rows (coalesced across threads) and accumulates omega split into
// two 7-bit limbs against x = (word >> k) & 0x00010001 = plus_k + 2^16 minus_k, so that one IMAD serves both masks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lab_field.cuh"

namespace lab {

constexpr int JL2_UNIT = 512;                      // coefficients per table
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

// Shared-memory byte address of table entry (group parity b, subset m) for this lane: m * 256 + b * 128 + lane * 4, formed by ONE
// byte permute: result byte 0 = byte 0 of `off` (= b * 128 + lane * 4 < 256), byte 1 = byte BYTE of the mask word, rest zero.
template <int BYTE>
__device__ __forceinline__ uint32_t jl2_addr(uint32_t x, uint32_t off) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(off), "n"(0x7704 | (BYTE << 4)));   // selectors: [3]=off.b3 (0) [2]=off.b3 (0) [1]=x.bBYTE [0]=off.b0
    return r;
}
__device__ __forceinline__ int jl2_lds(uint32_t saddr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// pi2: rows of this call's witness vectors, [ni][256][W] with W = ND / 16 words per row; S: the full witness [R][ND] (device);
// vector li of pi2 is witness vector i0 + li.  p: int64[256], accumulated with atomics (zero it first).
__global__ void __launch_bounds__(JL2_THREADS, 3) k_jl2(const uint32_t *__restrict__ pi2, const uint32_t *__restrict__ S, uint64_t ND, uint32_t W,
                                                         uint32_t i0, uint32_t units_per_vec, uint64_t total_units, unsigned long long *__restrict__ p) {
    extern __shared__ __align__(256) uint32_t tab[];            // [m][group parity][32 lanes]: 256 bytes per subset m
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t off0 = (uint32_t)lane * 4u, off1 = off0 + 128u;
    int acc[32];
#pragma unroll
    for (int r = 0; r < 32; r++) acc[r] = 0;
    for (uint64_t unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const uint32_t li = (uint32_t)(unit / units_per_vec), u = (uint32_t)(unit % units_per_vec);
        // ---- the lane's 16 coefficients (groups 2 lane, 2 lane + 1), canonical; past the end of the vector: zero ----
        uint32_t sv[16];
        {
            const uint64_t c0 = (uint64_t)u * JL2_UNIT + (uint64_t)lane * 16;
            const bool in = c0 < ND;                                    // ND is a multiple of 64: all 16 or none
            const uint4 *src = reinterpret_cast<const uint4 *>(S + (uint64_t)(i0 + li) * ND + (in ? c0 : 0));
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 v = in ? __ldg(src + q) : make_uint4(0, 0, 0, 0);
                sv[4 * q] = lab_canon(v.x); sv[4 * q + 1] = lab_canon(v.y); sv[4 * q + 2] = lab_canon(v.z); sv[4 * q + 3] = lab_canon(v.w);
            }
        }
        // the unit's words of rows 32 w .. 32 w + 15 are requested before the table is built.  A lane outside the row (ragged
        // last unit) reads word 0 of the unit instead and masks it: no predicated address arithmetic in the unrolled loads.
        const bool win = (uint64_t)u * (JL2_UNIT / 16) + lane < W;
        const uint32_t *row = pi2 + ((uint64_t)li * 256 + (uint64_t)w * 32) * W + (uint64_t)u * (JL2_UNIT / 16) + (win ? lane : 0);
        const uint32_t keep = win ? 0xFFFFFFFFu : 0u;
        uint32_t wd[16];
#pragma unroll
        for (int r = 0; r < 16; r++) wd[r] = __ldg(row + (uint32_t)r * W);
        __syncthreads();                                                // every warp is done with the previous table
        // ---- subset sums: warp w fills m = 32 w + k (k in Gray-code order: one add or subtract per entry) ----
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const uint32_t *s8 = sv + 8 * b;
            uint32_t val = ((w & 1) ? s8[5] : 0u) + ((w & 2) ? s8[6] : 0u) + ((w & 4) ? s8[7] : 0u);
            uint32_t *t = tab + (size_t)(w * 32) * 64 + b * 32 + lane;
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
        // ---- lookups: rows 32 w .. 32 w + 31, two halves of 16 rows ----
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t nx[16];
            if (half == 0) {                                            // second half of the rows: in flight during the first half's lookups
#pragma unroll
                for (int r = 0; r < 16; r++) nx[r] = __ldg(row + (uint32_t)(16 + r) * W);
            }
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const uint32_t x = wd[r] & keep;
                const int a0 = jl2_lds(tab_s + jl2_addr<0>(x, off0));   // plus mask, group 2 lane
                const int a1 = jl2_lds(tab_s + jl2_addr<1>(x, off1));   // plus mask, group 2 lane + 1
                const int m0 = jl2_lds(tab_s + jl2_addr<2>(x, off0));   // minus mask, group 2 lane
                const int m1 = jl2_lds(tab_s + jl2_addr<3>(x, off1));   // minus mask, group 2 lane + 1
                acc[16 * half + r] += (a0 - m0) + (a1 - m1);
            }
            if (half == 0) {
#pragma unroll
                for (int r = 0; r < 16; r++) wd[r] = nx[r];
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

// Small shapes (a default-size proof has 2 x 256 x 128 entries): the table machinery above would be all set-up.  One CTA per
// (vector, chunk of 1024 coefficients), thread = row; the chunk of s sits in 4 KB of static shared memory (every thread reads
// the same coefficient at the same time: a broadcast), the row's words come straight from global memory.
constexpr int JL2S_CH = 1024;
__global__ void __launch_bounds__(256) k_jl2_small(const uint32_t *__restrict__ pi2, const uint32_t *__restrict__ S, uint64_t ND, uint32_t W, uint32_t i0,
                                                   uint32_t chunks_per_vec, unsigned long long *__restrict__ p) {
    __shared__ uint32_t ss[JL2S_CH];
    const uint32_t li = blockIdx.x / chunks_per_vec, ch = blockIdx.x % chunks_per_vec;
    const uint64_t c0 = (uint64_t)ch * JL2S_CH;
    const uint32_t len = (uint32_t)min((uint64_t)JL2S_CH, ND - c0);            // multiple of 64
    for (uint32_t t = threadIdx.x; t < len; t += 256) ss[t] = lab_canon(S[(uint64_t)(i0 + li) * ND + c0 + t]);
    __syncthreads();
    const uint32_t *row = pi2 + ((uint64_t)li * 256 + threadIdx.x) * W + c0 / 16;
    int acc = 0;
    for (uint32_t wq = 0; wq < len / 16; wq++) {
        const uint32_t x = __ldg(row + wq);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int sv = (int)ss[wq * 16 + k];
            acc += ((x >> k) & 1u) ? sv : 0;
            acc -= ((x >> (16 + k)) & 1u) ? sv : 0;
        }
    }
    if (acc) atomicAdd(p + threadIdx.x, (unsigned long long)(long long)acc);
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

// The same for small shapes (a default-size proof has 16 words: 16 threads walking 256 rows each would be one long dependent
// chain): one WARP per word, lane = 8 of the 256 rows, the 32 packed sums reduced across the lanes with shuffles.
__global__ void __launch_bounds__(256) k_piT_omega2_warp(const uint32_t *__restrict__ pi2, const uint32_t *__restrict__ omega, uint64_t total_words, uint32_t W,
                                                         uint32_t *__restrict__ v) {
    __shared__ uint32_t som[2][256];
    {
        const uint32_t o = lab_canon(omega[threadIdx.x]);
        som[0][threadIdx.x] = o & 127u;
        som[1][threadIdx.x] = o >> 7;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t idx = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (idx >= total_words) return;
    const uint64_t i = idx / W, wq = idx % W;
    const uint32_t *col = pi2 + i * 256 * (uint64_t)W + wq;
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { lo[k] = 0; hi[k] = 0; }
    uint32_t x[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) x[jj] = __ldg(col + (uint64_t)(lane * 8 + jj) * W);
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
        const uint32_t ol = som[0][lane * 8 + jj], oh = som[1][lane * 8 + jj];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t t = (x[jj] >> k) & 0x00010001u;
            lo[k] += ol * t;
            hi[k] += oh * t;
        }
    }
    // each 16-bit half stays below 2^15 over all 256 rows, so packed words add without carries between halves
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[k] += __shfl_xor_sync(0xffffffffu, lo[k], o);
            hi[k] += __shfl_xor_sync(0xffffffffu, hi[k], o);
        }
    if (lane == 0) {
        uint32_t out[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t pos = (lo[k] & 0xFFFFu) + 128u * (hi[k] & 0xFFFFu);
            const uint32_t neg = (lo[k] >> 16) + 128u * (hi[k] >> 16);
            out[k] = lab_canon(pos + 257u * LABQ - neg);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(v + idx * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) dst[q] = make_uint4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    }
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
