// CRS-resident inner commitment on the tensor cores (tcgen05, sm_100a).
//
// When the CRS cache (lab_crs_cache_configure) holds A, the commitment t_i = A s_i (proofgen.rs:41-49) no longer runs
// ChaCha20 and is bound by its 2^41 complex multiply-accumulates (cfg 3).  On the CUDA cores that is 8.8e12 IMADs
// (0.47 s of FMA-pipe time); as int8 limb products on the 5th-generation tensor cores it is 3.5e13 MACs (16 ms at peak),
// so the kernel becomes a stream of the 137 GB of A through HBM.  Exactness: every product is an integer, the s32
// accumulators cannot overflow for K' <= 32768, and the epilogue reduces mod q = 2^13 - 1.
//
// Per transform slot j (32 independent contractions over F_{q^2}):
//   A_j  [2 * rows][K' = 2N] s8, K-major: for each block of 64 rows first the 64 "lo" rows (value & 127), then the 64 "hi"
//        rows (value >> 7); k = 2n + {0: re, 1: im}.  This IS the cache layout: K_A's producers write it (MODE 1).
//   B_j  [4 * R][K'] s8, K-major: four columns per witness vector i, (RE,lo) (RE,hi) (IM,lo) (IM,hi), with
//        B[(RE,l)][(n,re)] = s_re^l, B[(RE,l)][(n,im)] = -s_im^l, B[(IM,l)][(n,re)] = s_im^l, B[(IM,l)][(n,im)] = s_re^l
//   D = A_j B_j^T (s32 in TMEM, M = 128 = 64 rows x {lo, hi}, N' = 4R <= 256):
//        t_re = D_lo[RE,lo] + 2^7 (D_lo[RE,hi] + D_hi[RE,lo]) + 2^14 D_hi[RE,hi]   (2^14 = 2 mod q), same for im.
// One CTA per SM, persistent over (slot, row block) work items; warp 0 = TMA producer (4-stage ring of 128-byte-swizzled
// K-chunks), warp 1 = MMA issuer (tcgen05.mma.kind::i8, one thread), warps 2-5 = epilogue (tcgen05.ld, limb recombination,
// reduction mod q, coalesced store into slot planes); accumulators double-buffered in the 512 TMEM columns.
#pragma once
#include <cuda.h>
#include "lab_crsgen.cuh"

namespace lab {

constexpr int UM_STAGES = 4;
constexpr int UM_KC = 128;                 // K bytes per stage = one 128-byte swizzle row
constexpr int UM_THREADS = 192;            // producer warp, MMA warp, four epilogue warps
constexpr size_t UM_SMEM = (size_t)UM_STAGES * (128 * UM_KC + 256 * UM_KC) + 1024;

__device__ __forceinline__ uint32_t um_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void um_mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void um_mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nUM_W:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra UM_D;\nbra UM_W;\nUM_D:\n}" ::"r"(um_smem(b)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void um_mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(um_smem(b)) : "memory"); }
__device__ __forceinline__ void um_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(um_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void um_tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(um_smem(dst)), "l"(map),
                 "r"(c0), "r"(c1), "r"(um_smem(bar))
                 : "memory");
}
// K-major operand tile in SWIZZLE_128B shared memory: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t um_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// CRS side, cold: rows [row0, row0 + nrows) of A regenerated from the seed (ChaCha20 + warp transform) straight into int8
// limb planes -- no multiply-accumulate in this kernel, every warp generates, two polynomials (four ChaCha20 states per lane)
// per iteration.  A warp takes runs of 16 consecutive columns of one row, so that the two 16-bit stores per lane and polynomial
// fill whole 32-byte sectors in L2 before they reach HBM.
constexpr int GP_WARPS = 8;
#ifndef LAB_GP_MINB
#define LAB_GP_MINB 2
#endif
#ifndef LAB_GP_VAR
#define LAB_GP_VAR 21         /* measured, profiles/kbench_r2b_gen_variants.jsonl: split shuffles (+0.6 %), two polynomials = four ChaCha20 states per lane and
                                 iteration with the middle double rounds unrolled x2 (+2.9 % together; x8 of the one-polynomial form +2.3 %, x8 of this form
                                 overflows the instruction cache: -25 %); bit 1 (32-bit window check) measured equal */
#endif
template <int MINB, int VAR = LAB_GP_VAR>
__global__ void __launch_bounds__(32 * GP_WARPS, MINB) k_gen_planes(LabSeed seed, uint32_t N, uint64_t row0, uint64_t nrows, uint8_t *__restrict__ planes,
                                                                 uint32_t ntiles, uint32_t kpad) {
    __shared__ uint32_t tws[LAB_TWS_ROWS][32];
    __shared__ uint32_t hoist[GP_WARPS][16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w == 0) lab_warp_tw_to_smem(tws, lane);
    __syncthreads();
    const uint64_t runs_per_row = (N + 15) / 16, total_runs = nrows * runs_per_row;
    const uint64_t nwarps = (uint64_t)gridDim.x * GP_WARPS;
    LabWarpGen g;
    for (uint64_t run = (uint64_t)blockIdx.x * GP_WARPS + w; run < total_runs; run += nwarps) {
        const uint64_t row = run / runs_per_row;
        const uint32_t n0 = (uint32_t)(run % runs_per_row) * 16, n1 = min(n0 + 16u, N);
        uint8_t *q8 = planes + (((uint64_t)lane * ntiles + (row >> 6)) * 128 + (row & 63)) * kpad;
        uint32_t n = n0;
        if (VAR & 16) {                                                              // two polynomials per iteration
            for (; n + 1 < n1; n += 2) {
                const uint64_t ctr = ((row0 + row) * (uint64_t)N + n) * 64ull;          // structs.rs:55-72
                uint32_t re[2], im[2];
                lab_crs_poly_hat_warp_x2<VAR>(seed, g, hoist[w], tws, ctr, 0ull, lane, re, im);
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    *reinterpret_cast<uint16_t *>(q8 + 2 * (n + p)) = (uint16_t)((re[p] & 127u) | ((im[p] & 127u) << 8));
                    *reinterpret_cast<uint16_t *>(q8 + 64 * (uint64_t)kpad + 2 * (n + p)) = (uint16_t)((re[p] >> 7) | ((im[p] >> 7) << 8));
                }
            }
        }
        for (; n < n1; n++) {
            const uint64_t ctr = ((row0 + row) * (uint64_t)N + n) * 64ull;          // structs.rs:55-72
            uint32_t re, im;
            lab_crs_poly_hat_warp<VAR>(seed, g, hoist[w], tws, ctr, 0ull, lane, re, im, n == n0);
            *reinterpret_cast<uint16_t *>(q8 + 2 * n) = (uint16_t)((re & 127u) | ((im & 127u) << 8));
            *reinterpret_cast<uint16_t *>(q8 + 64 * (uint64_t)kpad + 2 * n) = (uint16_t)((re >> 7) | ((im >> 7) << 8));
        }
    }
}

// witness side: B planes from the n-major transformed witness What[(n * R + i) * 32 + j] (packed re | im << 16), vectors
// [i_base, i_base + ni); Bp[j][col][kpad], col = 4 * (i - i_base) + c; columns >= 4 * ni and k >= 2N stay zero (memset)
__global__ void __launch_bounds__(256) k_umma_build_b(const uint32_t *__restrict__ What, uint32_t N, uint32_t R, uint32_t i_base, uint32_t ni, uint32_t ncols,
                                                      uint32_t kpad, int8_t *__restrict__ Bp) {
    const size_t total = (size_t)N * ni * 32;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(idx & 31);
        const uint32_t i = (uint32_t)((idx >> 5) % ni);
        const uint32_t n = (uint32_t)((idx >> 5) / ni);
        const uint32_t v = What[((size_t)n * R + i_base + i) * 32 + j];
        const int re = (int)lab_re(v), im = (int)lab_im(v);
        const int rl = re & 127, rh = re >> 7, il = im & 127, ih = im >> 7;
        int8_t *base = Bp + ((size_t)j * ncols + 4 * i) * kpad + 2 * n;
        auto put = [&](int col, int a, int b) { *reinterpret_cast<uint16_t *>(base + (size_t)col * kpad) = (uint16_t)((uint8_t)(int8_t)a | ((uint16_t)(uint8_t)(int8_t)b << 8)); };
        put(0, rl, -il);
        put(1, rh, -ih);
        put(2, il, rl);
        put(3, ih, rh);
    }
}

// Th[j][v][rows_pad] packed slots (re | im << 16), v = vector index within the pass
__global__ void __launch_bounds__(UM_THREADS, 1) k_umma_commit(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                               uint32_t ntiles /* row blocks of 64 */, uint32_t chunk0, uint32_t nchunks /* K range in 128-byte chunks */,
                                                               uint32_t ncols /* N' */, uint32_t nvec, uint32_t rows_pad, uint32_t *__restrict__ Th) {
    extern __shared__ __align__(1024) uint8_t um_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)um_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                   // [stages][128][128]
    uint8_t *sB = smem + UM_STAGES * 128 * UM_KC;         // [stages][256][128]
    __shared__ uint64_t full[UM_STAGES], empty[UM_STAGES], tfull[2], tempty[2];
    __shared__ uint32_t tmem_base;
    __shared__ uint32_t xch[64][17];                      // hi-limb rows -> lo-limb threads, 16 columns at a time
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem(&tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; s++) { um_mbar_init(&full[s], 1); um_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { um_mbar_init(&tfull[b], 1); um_mbar_init(&tempty[b], 64); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tmem_base;
    const uint32_t nwork = 32u * ntiles;
    const uint32_t stage_tx = 128 * UM_KC + ncols * UM_KC;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x) {
                const uint32_t j = w / ntiles, tile = w % ntiles;
                const int rowA = (int)((j * ntiles + tile) * 128u), rowB = (int)(j * ncols);
                for (uint32_t c = 0; c < nchunks; c++, it++) {
                    const uint32_t s = it % UM_STAGES;
                    if (it >= UM_STAGES) um_mbar_wait(&empty[s], ((it / UM_STAGES) - 1) & 1);
                    um_expect_tx(&full[s], stage_tx);
                    um_tma_2d(sA + s * 128 * UM_KC, &mapA, (int)((chunk0 + c) * UM_KC), rowA, &full[s]);
                    um_tma_2d(sB + s * 256 * UM_KC, &mapB, (int)((chunk0 + c) * UM_KC), rowB, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            // D = s32, A = B = signed 8 bit, both K-major, N = ncols, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((ncols >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t it = 0, wi = 0;
            for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x, wi++) {
                const uint32_t buf = wi & 1;
                if (wi >= 2) um_mbar_wait(&tempty[buf], ((wi >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t tacc = tm + buf * 256u;
                for (uint32_t c = 0; c < nchunks; c++, it++) {
                    const uint32_t s = it % UM_STAGES;
                    um_mbar_wait(&full[s], (it / UM_STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint64_t da = um_desc(um_smem(sA + s * 128 * UM_KC)), db = um_desc(um_smem(sB + s * 256 * UM_KC));
#pragma unroll
                    for (int k = 0; k < UM_KC / 32; k++) {
                        const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tacc),
                            "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(acc)
                            : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(um_smem(&empty[s])) : "memory");   // smem slot reusable
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(um_smem(&tfull[buf])) : "memory");       // accumulator complete
            }
        }
    } else {
        // ---------------- epilogue: warps 2..5 own TMEM lanes 32 * (warp % 4) ... ----------------
        const int quarter = warp & 3;                    // lanes [32q, 32q + 32): q = 0,1 -> lo rows 0..63, q = 2,3 -> hi rows 0..63
        const bool hi = quarter >= 2;
        const int r = (quarter & 1) * 32 + lane;         // row within the block of 64
        constexpr uint32_t OFF = LABQ * 131072u;         // multiple of q above any |accumulator| (K' <= 32768: 32768 * 127^2 < 2^29.1)
        uint32_t wi = 0;
        for (uint32_t w = blockIdx.x; w < nwork; w += gridDim.x, wi++) {
            const uint32_t j = w / ntiles, tile = w % ntiles, buf = wi & 1;
            um_mbar_wait(&tfull[buf], (wi >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (uint32_t c0 = 0; c0 < ncols; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tm + buf * 256u + ((uint32_t)(quarter * 32) << 16) + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                               "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
                for (int q = 0; q < 16; q++) v[q] = lab_canon(v[q] + OFF);          // signed accumulator -> residue in [0, q)
                if (hi) {
#pragma unroll
                    for (int q = 0; q < 16; q++) xch[r][q] = v[q];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (!hi) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {                                    // four vectors per 16 columns
                        const uint32_t vec = c0 / 4 + u;
                        // lo * lo + 2^7 (lo * hi + hi * lo) + 2^14 hi * hi, 2^14 = 2 (mod q)
                        const uint32_t re = lab_canon(v[4 * u] + 128u * (v[4 * u + 1] + xch[r][4 * u]) + 2u * xch[r][4 * u + 1]);
                        const uint32_t im = lab_canon(v[4 * u + 2] + 128u * (v[4 * u + 3] + xch[r][4 * u + 2]) + 2u * xch[r][4 * u + 3]);
                        if (vec < nvec) Th[((size_t)j * nvec + vec) * rows_pad + (size_t)tile * 64 + r] = lab_pack(re, im);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            if (!hi) um_mbar_arrive(&tempty[buf]);       // 64 lo-row threads: accumulator buffer may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

// slot planes -> T: CTA = (32 rows, one vector); sums the nseg K-segments (s32 accumulators hold at most 32768 bytes of K each),
// transposes [32 slots][32 rows] through shared memory, inverse transform per row
__global__ void __launch_bounds__(256) k_umma_finish(const uint32_t *__restrict__ Th, uint32_t nseg, size_t seg_stride, uint32_t nvec, uint32_t rows_pad,
                                                     uint64_t nrows, uint32_t i_base, uint32_t *__restrict__ T, uint64_t t_stride, uint64_t t_row_off) {
    __shared__ uint32_t tile[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t vec = blockIdx.y;
    const uint64_t row0 = (uint64_t)blockIdx.x * 32;
    for (int s = w; s < 32; s += 8) {
        uint32_t re = 0, im = 0;
        for (uint32_t g = 0; g < nseg; g++) {
            const uint32_t v = Th[g * seg_stride + ((size_t)s * nvec + vec) * rows_pad + row0 + lane];
            re += lab_re(v); im += lab_im(v);
        }
        tile[s][lane] = lab_pack(lab_canon(re), lab_canon(im));
    }
    __syncthreads();
    const LabWarpTw tw = lab_warp_tw(lane);
    for (int rr = w; rr < 32; rr += 8) {
        const uint64_t row = row0 + rr;
        uint32_t v = tile[lane][rr];
        uint32_t re = lab_re(v), im = lab_im(v);
        lab_ntt32_inv_warp(re, im, tw, lane);
        if (row < nrows) {
            uint32_t *dst = T + ((size_t)(i_base + vec) * t_stride + t_row_off + row) * 64;
            dst[lane] = re;
            dst[lane + 32] = im;
        }
    }
}

}  // namespace lab
