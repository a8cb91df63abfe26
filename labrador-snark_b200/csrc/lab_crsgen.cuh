// One CRS polynomial per warp, straight into the transform domain -- the producer step shared by the CRS-regenerating
// kernels (k_gen_planes, k_crs_matvec): trimmed ChaCha20 for coefficients lane and lane + 32 (lab_chacha.cuh), rand-0.8.5
// sampling from keystream word 3 (lab_sample_w3), generic path for draws it does not decide and for polynomials that
// straddle a 2^32 boundary of seed + counter, warp transform with the constants in shared memory.  The hoisted part of the
// first double round lives in a per-warp shared-memory slot (15 words) and is recomputed when the high part of seed + counter
// changes, so a producer thread carries its ChaCha20 states and little else: two states in 72 registers (one polynomial per
// call), four in 128 (lab_crs_poly_hat_warp_x2, what k_gen_planes runs: two CTAs of 256 threads per SM).
#pragma once
#include "lab_chacha.cuh"
#include "lab_ntt.cuh"

struct LabWarpGen {
    uint64_t tag_lo = ~0ull, tag_hi = ~0ull;      // which counter range the warp's hoist slot is valid for
};

__device__ __forceinline__ void lab_hoist_store(uint32_t *q, const LabHoist &hh) {
    q[0] = hh.k3; q[1] = hh.P0; q[2] = hh.P1; q[3] = hh.Q0; q[4] = hh.A5; q[5] = hh.A10; q[6] = hh.Q1; q[7] = hh.Q2;
    q[8] = hh.A6; q[9] = hh.A2; q[10] = hh.A8; q[11] = hh.A13; q[12] = hh.A4; q[13] = hh.A9; q[14] = hh.A14;
}
__device__ __forceinline__ void lab_hoist_load(const uint32_t *q, LabHoist &h) {
    h.k3 = q[0]; h.P0 = q[1]; h.P1 = q[2]; h.Q0 = q[3]; h.A5 = q[4]; h.A10 = q[5]; h.Q1 = q[6]; h.Q2 = q[7];
    h.A6 = q[8]; h.A2 = q[9]; h.A8 = q[10]; h.A13 = q[11]; h.A4 = q[12]; h.A9 = q[13]; h.A14 = q[14];
}

// polynomial whose coefficient 0 sits at counter (chi:clo); all lanes of the warp call with the same arguments.
// hoist_slot: 16 words of shared memory owned by this warp; tws: lab_warp_tw_to_smem table.  Returns the packed slots.
// VAR bits 2-3: log2 of the unroll factor of the double-round loop; VAR bit 0: split shuffles in the transform; VAR bit 1: the caller walks consecutive polynomials (clo advances by 64) and has
// called this function for the previous one, so the hoist slot can only have become stale if the low 32 bits of seed + counter
// wrapped -- one 32-bit compare instead of the 64-bit tag comparison (`first` = first polynomial of a run: full check)
template <int VAR = 0>
__device__ __forceinline__ void lab_crs_poly_hat_warp(const LabSeed &seed, LabWarpGen &g, uint32_t *hoist_slot, const uint32_t (*tws)[32], uint64_t clo,
                                                      uint64_t chi, int lane, uint32_t &re, uint32_t &im, bool first = true) {
    uint32_t lo32;
    if ((VAR & 2) && !first && (lo32 = (uint32_t)seed.limb[0] + (uint32_t)clo) >= 64u) {
        // same 2^32 window as the previous polynomial: the slot is current
    } else {
        const uint64_t s0 = seed.limb[0] + clo;
        const uint64_t ntag = ((uint64_t)(s0 < clo) << 32) | (s0 >> 32);
        if (ntag != g.tag_lo || chi != g.tag_hi) {                   // warp-uniform, once per 2^32 counters
            LabHoist hh;
            lab_hoist_compute(seed, clo, chi, hh);
            __syncwarp();
            if (lane == 0) lab_hoist_store(hoist_slot, hh);
            __syncwarp();
            g.tag_lo = ntag;
            g.tag_hi = chi;
        }
        lo32 = (uint32_t)s0;
    }
    const bool straddle = lo32 > 0xFFFFFFFFu - 63u;
    LabHoist h;
    lab_hoist_load(hoist_slot, h);
    const uint32_t k7[2] = {lab_bswap32(lo32 + (uint32_t)lane), lab_bswap32(lo32 + (uint32_t)lane + 32u)};
    uint32_t w3[2], c[2];
    lab_chacha_w3<2, 0u, 1 << ((VAR >> 2) & 3)>(seed, h, k7, w3);
    uint32_t slow = straddle ? 3u : 0u;
    slow |= lab_sample_w3(w3[0], c[0]) ? 0u : 1u;
    slow |= lab_sample_w3(w3[1], c[1]) ? 0u : 2u;
    if (slow) {
        if (slow & 1u) {
            const uint64_t l0 = clo + (uint64_t)lane;
            c[0] = lab_crs_coeff_generic(seed, l0, chi + (l0 < clo), 0u);
        }
        if (slow & 2u) {
            const uint64_t l1 = clo + (uint64_t)lane + 32u;
            c[1] = lab_crs_coeff_generic(seed, l1, chi + (l1 < clo), 0u);
        }
    }
    re = c[0]; im = c[1];
    lab_ntt32_fwd_warp_smem<(VAR & 1) != 0>(re, im, tws, lane, seed.one);
}

// The polynomials at counters (chi:clo) and (chi:clo) + 64 together: four interleaved ChaCha20 states per lane (the round
// function alone runs 6 % faster with four independent blocks per thread than with two, profiles/kbench_r1_pipe_ceilings.jsonl)
// and two interleaved transforms.  VAR as above.
template <int VAR = 0>
__device__ __forceinline__ void lab_crs_poly_hat_warp_x2(const LabSeed &seed, LabWarpGen &g, uint32_t *hoist_slot, const uint32_t (*tws)[32], uint64_t clo,
                                                         uint64_t chi, int lane, uint32_t (&re)[2], uint32_t (&im)[2]) {
    const uint64_t s0 = seed.limb[0] + clo;
    const uint64_t ntag = ((uint64_t)(s0 < clo) << 32) | (s0 >> 32);
    if (ntag != g.tag_lo || chi != g.tag_hi) {                   // warp-uniform, once per 2^32 counters
        LabHoist hh;
        lab_hoist_compute(seed, clo, chi, hh);
        __syncwarp();
        if (lane == 0) lab_hoist_store(hoist_slot, hh);
        __syncwarp();
        g.tag_lo = ntag;
        g.tag_hi = chi;
    }
    const uint32_t lo32 = (uint32_t)s0;
    const bool straddle = lo32 > 0xFFFFFFFFu - 127u;             // some of the 128 counters lie in the next 2^32 window: generic path
    LabHoist h;
    lab_hoist_load(hoist_slot, h);
    uint32_t k7[4], w3[4], c[4];
#pragma unroll
    for (int b = 0; b < 4; b++) k7[b] = lab_bswap32(lo32 + (uint32_t)lane + 32u * b);
    lab_chacha_w3<4, 0u, 1 << ((VAR >> 2) & 3)>(seed, h, k7, w3);
    uint32_t slow = straddle ? 15u : 0u;
#pragma unroll
    for (int b = 0; b < 4; b++) slow |= lab_sample_w3(w3[b], c[b]) ? 0u : 1u << b;
    if (slow) {
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (slow >> b & 1u) {
                const uint64_t l = clo + (uint64_t)lane + 32u * b;
                c[b] = lab_crs_coeff_generic(seed, l, chi + (l < clo), 0u);
            }
    }
    re[0] = c[0]; im[0] = c[1]; re[1] = c[2]; im[1] = c[3];
    lab_ntt32_fwd_warp_smem_x2<(VAR & 1) != 0>(re, im, tws, lane, seed.one);
}
