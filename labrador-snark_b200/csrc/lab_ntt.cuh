// 32-point transforms over F_{Q^2}: R_q = F_Q[X]/(X^64+1) ~ F_{Q^2}[X]/(X^32 - i) ~ F_{Q^2}^32.
// Two mappings of the same network (tables and math in tools/gen_ntt.py):
//   * lab_ntt32_{fwd,inv}_regs  -- one polynomial per LANE, 32 complex values in registers,
//                                  straight-line generated code (batch NTT / polymul kernels)
//   * lab_ntt32_{fwd,inv}_warp  -- one polynomial per WARP, lane j holds complex slot j,
//                                  butterflies by __shfl_xor (CRS-fused commitment kernels)
// This replaces concrete-ntt's Plan32::negacyclic_polymul (algebraic.rs:396): any exact
// negacyclic product mod Q is bit-identical to it.
#pragma once
#include "lab_field.cuh"
#include "lab_ntt_gen.cuh"

#ifdef __CUDACC__
static __constant__ uint32_t LAB_TW_FWD[32] = LAB_TW_FWD_INIT;
static __constant__ uint32_t LAB_TW_INV[32] = LAB_TW_INV_INIT;

// per-lane multipliers: stage s (len = 16 >> s) uses tree node (1 << s) + (lane >> (5 - s)).
// Lower lanes of a butterfly multiply by 1 so that the code is branch-free.  The forward constants are kept
// unpacked (re, im, -im) together with the butterfly sign and offset, so that a stage costs only 6 ALU-pipe
// instructions (4 fold shifts + 2 unpacks); everything else is IMAD on the FMA pipe.
// Bounds of a forward stage (asserted on the host in tests/test_host.py): values enter < Q + 16, the twiddle parts are
// <= Q, so a product sum is < 2^27.01 and ONE fold leaves it <= 24595 < 4Q -- small enough for the 16-bit halves of the
// shuffled word and for the upper lanes' lo - t + 4Q to stay positive; the sum of a butterfly is < 2^16, and one more fold
// brings it back below Q + 16.
struct LabWarpTw {
    uint32_t fr[5], fi[5], nfi[5];   // forward: upper lanes zeta^(e/2), lower lanes 1; nfi = Q - fi
    uint32_t sgn[5], off[5];         // forward butterfly: out = p * sgn + (recv + off); upper: (-1, 4Q), lower: (1, 0)
    uint32_t g[5];                   // inverse: upper lanes zeta^-(e/2), lower lanes 1; last level also carries 2^8 = 32^-1
};

__device__ __forceinline__ LabWarpTw lab_warp_tw(int lane) {
    LabWarpTw t;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int len = 16 >> s;
        const bool upper = (lane & len) != 0;
        const int node = (1 << s) + (lane >> (5 - s));
        const uint32_t f = upper ? LAB_TW_FWD[node] : 1u;
        t.fr[s] = lab_re(f);
        t.fi[s] = lab_im(f);
        t.nfi[s] = LABQ - lab_im(f);
        t.sgn[s] = upper ? 0xFFFFFFFFu : 1u;
        t.off[s] = upper ? 4u * LABQ : 0u;
        uint32_t g = upper ? LAB_TW_INV[node] : 1u;
        if (s == 0) {   // len == 16 is the LAST inverse level: fold in 32^-1 = 256
            uint32_t gr = lab_canon(lab_re(g) * 256u), gi = lab_canon(lab_im(g) * 256u);
            g = lab_pack(gr, gi);
        }
        t.g[s] = g;
    }
    return t;
}

// lane j holds g_j = f_j + i f_{j+32} (residues < 2Q); returns slot j = f(zeta^{e_j}), canonical.
// `one` == 1 at run time (LabSeed::one): additions written as x * one + y issue on the FMA pipe.
template <bool SPLIT = false>
__device__ __forceinline__ void lab_ntt32_fwd_warp(uint32_t &re, uint32_t &im, const LabWarpTw &tw, int lane, uint32_t one = 1u) {
    (void)lane;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int len = 16 >> s;
        uint32_t pr = re * tw.fr[s] + im * tw.nfi[s];                        // < 2^27.01
        uint32_t pi = re * tw.fi[s] + im * tw.fr[s];
        pr = lab_fold(pr);                                                   // <= 24595 < 4Q
        pi = lab_fold(pi);
        // lower: lo + t ; upper: lo - t + 4Q  (t is the upper lane's product, lo the lower lane's value)
        if (SPLIT) {                                                         // two shuffles, no pack / unpack (see lab_ntt32_fwd_warp_smem)
            const uint32_t rr = __shfl_xor_sync(0xffffffffu, pr, len), ri = __shfl_xor_sync(0xffffffffu, pi, len);
            re = lab_fold(pr * tw.sgn[s] + (rr * one + tw.off[s]));          // < Q + 16
            im = lab_fold(pi * tw.sgn[s] + (ri * one + tw.off[s]));
        } else {
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, pi * 65536u + pr, len);
            re = lab_fold(pr * tw.sgn[s] + (lab_re(recv) * one + tw.off[s]));
            im = lab_fold(pi * tw.sgn[s] + (lab_im(recv) * one + tw.off[s]));
        }
    }
    re = lab_csub(re);
    im = lab_csub(im);
}

// The forward constants of all 32 lanes as a table in shared memory (row 5 * kind + stage; kind = fr, fi, nfi, sgn, off):
// the CRS producer warps of K_A read them with LDS per stage instead of holding 25 registers across the ChaCha20 rounds.
constexpr int LAB_TWS_ROWS = 25;
__device__ __forceinline__ void lab_warp_tw_to_smem(uint32_t (*tws)[32], int lane) {
    const LabWarpTw t = lab_warp_tw(lane);
#pragma unroll
    for (int s = 0; s < 5; s++) {
        tws[s][lane] = t.fr[s];
        tws[5 + s][lane] = t.fi[s];
        tws[10 + s][lane] = t.nfi[s];
        tws[15 + s][lane] = t.sgn[s];
        tws[20 + s][lane] = t.off[s];
    }
}
// SPLIT: re and im travel in two shuffles instead of one packed word: two unpack operations less on the ALU pipe and one
// pack multiply less on the FMA pipe per stage, for one more SHFL
template <bool SPLIT = false>
__device__ __forceinline__ void lab_ntt32_fwd_warp_smem(uint32_t &re, uint32_t &im, const uint32_t (*tws)[32], int lane, uint32_t one) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int len = 16 >> s;
        const uint32_t fr = tws[s][lane], fi = tws[5 + s][lane], nfi = tws[10 + s][lane], sgn = tws[15 + s][lane], off = tws[20 + s][lane];
        uint32_t pr = re * fr + im * nfi;
        uint32_t pi = re * fi + im * fr;
        pr = lab_fold(pr);
        pi = lab_fold(pi);
        if (SPLIT) {
            const uint32_t rr = __shfl_xor_sync(0xffffffffu, pr, len), ri = __shfl_xor_sync(0xffffffffu, pi, len);
            re = lab_fold(pr * sgn + (rr * one + off));
            im = lab_fold(pi * sgn + (ri * one + off));
        } else {
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, pi * 65536u + pr, len);
            re = lab_fold(pr * sgn + (lab_re(recv) * one + off));
            im = lab_fold(pi * sgn + (lab_im(recv) * one + off));
        }
    }
    re = lab_csub(re);
    im = lab_csub(im);
}

// two polynomials at once (independent butterflies interleaved: twice the instruction-level parallelism around each shuffle)
template <bool SPLIT = true>
__device__ __forceinline__ void lab_ntt32_fwd_warp_smem_x2(uint32_t (&re)[2], uint32_t (&im)[2], const uint32_t (*tws)[32], int lane, uint32_t one) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int len = 16 >> s;
        const uint32_t fr = tws[s][lane], fi = tws[5 + s][lane], nfi = tws[10 + s][lane], sgn = tws[15 + s][lane], off = tws[20 + s][lane];
        uint32_t pr[2], pi[2];
#pragma unroll
        for (int p = 0; p < 2; p++) {
            pr[p] = lab_fold(re[p] * fr + im[p] * nfi);
            pi[p] = lab_fold(re[p] * fi + im[p] * fr);
        }
        if (SPLIT) {
            uint32_t rr[2], ri[2];
#pragma unroll
            for (int p = 0; p < 2; p++) { rr[p] = __shfl_xor_sync(0xffffffffu, pr[p], len); ri[p] = __shfl_xor_sync(0xffffffffu, pi[p], len); }
#pragma unroll
            for (int p = 0; p < 2; p++) {
                re[p] = lab_fold(pr[p] * sgn + (rr[p] * one + off));
                im[p] = lab_fold(pi[p] * sgn + (ri[p] * one + off));
            }
        } else {
            uint32_t recv[2];
#pragma unroll
            for (int p = 0; p < 2; p++) recv[p] = __shfl_xor_sync(0xffffffffu, pi[p] * 65536u + pr[p], len);
#pragma unroll
            for (int p = 0; p < 2; p++) {
                re[p] = lab_fold(pr[p] * sgn + (lab_re(recv[p]) * one + off));
                im[p] = lab_fold(pi[p] * sgn + (lab_im(recv[p]) * one + off));
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; p++) { re[p] = lab_csub(re[p]); im[p] = lab_csub(im[p]); }
}

// lane j holds slot j (residues < 2Q); returns g_j = f_j + i f_{j+32}, canonical, scaled by 1/32.
__device__ __forceinline__ void lab_ntt32_inv_warp(uint32_t &re, uint32_t &im, const LabWarpTw &tw, int lane) {
#pragma unroll
    for (int s = 4; s >= 0; s--) {
        const int len = 16 >> s;
        const bool upper = (lane & len) != 0;
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, lab_pack(re, im), len);
        const uint32_t rr = lab_re(recv), ri = lab_im(recv);
        // lower: u + v ; upper: (u - v) * zeta^-(e/2)
        uint32_t xr = upper ? rr + 2u * LABQ - re : re + rr;
        uint32_t xi = upper ? ri + 2u * LABQ - im : im + ri;
        xr = lab_fold(xr);
        xi = lab_fold(xi);
        lab_cmul(xr, xi, lab_re(tw.g[s]), lab_im(tw.g[s]), re, im);          // < 2Q
    }
    re = lab_csub(re);
    im = lab_csub(im);
}
#endif
