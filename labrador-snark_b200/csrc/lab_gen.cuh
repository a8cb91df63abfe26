// Device-side input and challenge generation (f2 challenge generation, f4 witness generation).
// Same seeded definitions as labrador_b200/synth.py and the oracle (SplitMix64 counter PRG, one stream per tensor),
// so every output is compared bit for bit in tests/test_gpu_parity.py.
//   k_challenge_polys   Verifier::fetch_challenge (verification.rs:460-489): coefficients drawn without replacement
//                       from {0 x23, 1 x31, 2 x10}, non-zero ones negated with probability 1/2
//                       (generate_polynomial_picky, util.rs:83-104), resampled while the 1000-sample operator-norm
//                       estimate exceeds T = 15 (operator_norm, util.rs:227-246) -- 1000 negacyclic products per
//                       candidate, done here as warp transforms
//   k_witness_*         generate_witness (proofgen.rs:460-518): uniform coefficients, then floor-halving of randomly
//                       picked polynomials (reduce_polynomial, util.rs:37-51) until the squared norm is <= beta^2
#pragma once
#include "lab_ntt.cuh"

namespace lab {

__device__ __forceinline__ uint64_t gen_prg_u64(uint64_t base, uint64_t idx) {
    uint64_t z = base + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t gen_prg_base(uint64_t seed, uint64_t stream) { return seed + stream * 0xD1342543DE82EF95ull; }

constexpr int CH_WARPS = 16;          // warps per challenge polynomial
constexpr int CH_SAMPLES = 1000;      // util.rs:234
// one CTA per challenge index; out[idx][64] canonical
__global__ void __launch_bounds__(32 * CH_WARPS) k_challenge_polys(uint64_t seed, uint32_t first_idx, uint32_t *__restrict__ out, uint32_t *__restrict__ n_candidates) {
    __shared__ uint32_t cand[64];
    __shared__ int exceeded;
    const uint64_t idx = first_idx + blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t bs = gen_prg_base(seed, 10 + (idx << 8)), bo = gen_prg_base(seed, 11 + (idx << 8));
    const LabWarpTw tw = lab_warp_tw(lane);
    uint64_t draw = 0, od = 0;                     // stream positions (draw is only used by thread 0)
    uint32_t tries = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            // dist.pop(ri) on the sorted multiset: only the counts matter
            int n0 = 23, n1 = 31, n2 = 10;
            for (int d = 0; d < 64; d++) {
                const int len = n0 + n1 + n2;
                const int ri = (int)__umul64hi(gen_prg_u64(bs, draw++), (uint64_t)len);
                uint32_t coeff;
                if (ri < n0) { coeff = 0; n0--; }
                else if (ri < n0 + n1) { coeff = 1; n1--; }
                else { coeff = 2; n2--; }
                const int sgn = (int)(gen_prg_u64(bs, draw++) >> 63);
                cand[d] = (coeff > 0 && sgn) ? LABQ - coeff : coeff;
            }
            exceeded = 0;
        }
        __syncthreads();
        uint32_t cre = cand[lane], cim = cand[lane + 32];
        lab_ntt32_fwd_warp(cre, cim, tw, lane);
        bool mine = false;
        for (int s = w; s < CH_SAMPLES; s += CH_WARPS) {
            const uint32_t r0 = (uint32_t)__umul64hi(gen_prg_u64(bo, od + (uint64_t)s * 64 + lane), (uint64_t)LABQ);
            const uint32_t r1 = (uint32_t)__umul64hi(gen_prg_u64(bo, od + (uint64_t)s * 64 + lane + 32), (uint64_t)LABQ);
            uint32_t re = r0, im = r1;
            lab_ntt32_fwd_warp(re, im, tw, lane);
            uint32_t pr, pi;
            lab_cmul(cre, cim, re, im, pr, pi);
            lab_ntt32_inv_warp(pr, pi, tw, lane);                      // canonical coefficients lane, lane + 32 of c * r
            unsigned long long a = (unsigned long long)pr * pr + (unsigned long long)pi * pi;
            unsigned long long b = (unsigned long long)r0 * r0 + (unsigned long long)r1 * r1;
#pragma unroll
            for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
            // the reference's f64 ratio (util.rs:238-240); both sums are below 2^53, sqrt and divide are IEEE-exact
            const double ratio = sqrt((double)a) / sqrt((double)b);
            mine |= ratio > 15.0;
        }
        if (mine && lane == 0) exceeded = 1;
        od += (uint64_t)CH_SAMPLES * 64;
        tries++;
        __syncthreads();
        const bool rej = exceeded != 0;
        __syncthreads();
        if (!rej) break;
    }
    if (threadIdx.x < 64) out[(size_t)blockIdx.x * 64 + threadIdx.x] = cand[threadIdx.x];
    if (threadIdx.x == 0 && n_candidates) n_candidates[blockIdx.x] = tries;
}

// ---- generate_witness ----
constexpr int WIT_LEVELS = 14;        // a 13-bit coefficient is zero after 13 halvings
// uniform coefficients (stream 1) and, per polynomial, its squared norm after h = 0..13 floor-halvings
__global__ void __launch_bounds__(256) k_witness_uniform(uint64_t seed, size_t n_polys, uint32_t *__restrict__ S, unsigned long long *__restrict__ normtab) {
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= n_polys) return;
    const uint64_t base = gen_prg_base(seed, 1);
    const uint32_t c0 = (uint32_t)__umul64hi(gen_prg_u64(base, wid * 64 + lane), (uint64_t)LABQ);
    const uint32_t c1 = (uint32_t)__umul64hi(gen_prg_u64(base, wid * 64 + lane + 32), (uint64_t)LABQ);
    S[wid * 64 + lane] = c0;
    S[wid * 64 + lane + 32] = c1;
#pragma unroll 1
    for (int h = 0; h < WIT_LEVELS; h++) {
        unsigned long long a = (unsigned long long)(c0 >> h) * (c0 >> h) + (unsigned long long)(c1 >> h) * (c1 >> h);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) normtab[wid * WIT_LEVELS + h] = a;
    }
}
// the sequential part (one thread): picks (n, i) from stream 2 until the running norm is <= bound; halvings[p] counts
// how often polynomial p = i * N + n was picked.  Each step is two PRG draws and one table lookup.
__global__ void k_witness_pick(uint64_t seed, uint64_t N, uint64_t R, unsigned long long bound, const unsigned long long *__restrict__ normtab,
                               uint32_t *__restrict__ halvings, unsigned long long *__restrict__ info /* [0] final norm, [1] draws */) {
    if (blockIdx.x || threadIdx.x) return;
    const uint64_t base = gen_prg_base(seed, 2);
    // the starting norm: sum of level-0 entries (u64 is exact: N * R * 64 * q^2 < 2^64 for every supported shape)
    unsigned long long norm = 0;
    for (uint64_t p = 0; p < N * R; p++) norm += normtab[p * WIT_LEVELS];
    uint64_t draw = 0;
    while (norm > bound) {
        const uint64_t n = __umul64hi(gen_prg_u64(base, draw), N);
        const uint64_t i = __umul64hi(gen_prg_u64(base, draw + 1), R);
        draw += 2;
        const uint64_t p = i * N + n;
        const uint32_t h = halvings[p];
        if (h + 1 < WIT_LEVELS) {         // level 13 is the zero polynomial: further picks change nothing
            norm -= normtab[p * WIT_LEVELS + h] - normtab[p * WIT_LEVELS + h + 1];
            halvings[p] = h + 1;
        }
    }
    info[0] = norm;
    info[1] = draw;
}
__global__ void __launch_bounds__(256) k_witness_apply(const uint32_t *__restrict__ halvings, size_t n_coeffs, uint32_t *__restrict__ S) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n_coeffs; idx += stride) {
        const uint32_t h = halvings[idx >> 6];
        S[idx] = h >= 32 ? 0u : S[idx] >> h;
    }
}
// symmetric uniform a_ij (stream 3, entry (i, j), i <= j, at PRG index (i * R + j) * 64 + d; structs.rs:289-318)
__global__ void k_statement_a(uint64_t seed, uint32_t R, uint32_t *__restrict__ a) {
    const uint32_t i = blockIdx.x, j = blockIdx.y, d = threadIdx.x;
    const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
    a[((size_t)i * R + j) * 64 + d] = (uint32_t)__umul64hi(gen_prg_u64(gen_prg_base(seed, 3), ((uint64_t)lo * R + hi) * 64 + d), (uint64_t)LABQ);
}

}  // namespace lab
