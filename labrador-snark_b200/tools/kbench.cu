// Variant sweep for the CRS-regenerating kernels on one B200 (developer tool, not part of the product library).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../csrc -o bin/kbench kbench.cu
// Sweeps the FMA-pipe rotation mask (lab_chacha.cuh) of k_crs_expand and k_commit_inner, checks every variant against
// the untrimmed generic ChaCha20 path (bit-exact) and prints one JSON line per measurement.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "lab_kernels.cuh"
#include "lab_umma.cuh"

using namespace lab;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

static LabSeed mk_seed() {
    LabSeed s;
    uint8_t b[32];
    for (int i = 0; i < 32; i++) b[i] = (uint8_t)i;
    for (int l = 0; l < 4; l++) {
        uint64_t v = 0;
        for (int k = 0; k < 8; k++) v = (v << 8) | b[(3 - l) * 8 + k];
        s.limb[l] = v;
    }
    s.one = 1u; s.p16 = 1u << 16; s.p12 = 1u << 12; s.p8 = 1u << 8; s.p7 = 1u << 7;
    s.pad[0] = s.pad[1] = s.pad[2] = 0;
    return s;
}

// every coefficient through the generic (untrimmed, unhoisted) path
__global__ void k_expand_generic(LabSeed seed, uint64_t start_lo, size_t n, uint32_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lab_crs_coeff_slow(seed, start_lo + i, 0ull, 0u);
}
__global__ void k_fill_hat(uint32_t *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z ^= z >> 27;
        p[i] = (uint32_t)(z % 8191u) | ((uint32_t)((z >> 32) % 8191u) << 16);
    }
}
__global__ void k_checksum(const uint32_t *p, size_t n, unsigned long long *out) {
    unsigned long long a = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a += (unsigned long long)p[i] * (2 * (i % 1000003) + 1);
    atomicAdd(out, a);
}
// pipe microbenchmarks: chains of dependent ops, 8 independent chains per thread
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(uint32_t *out, int iters, uint32_t pw, uint32_t one) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = threadIdx.x * 8 + j + blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) v[j] = lab_rotl(v[j], 7) ^ pw;                 // SHF + LOP3 (ALU, 2 instr)
            if (MODE == 1) v[j] = v[j] * pw + one;                        // IMAD (1 instr)
            if (MODE == 2) v[j] = lab_rotl_fma(v[j], pw);                 // IMAD + IMAD.HI (2 instr)
            if (MODE == 3) { asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(v[j]) : "r"(v[j]), "r"(pw), "r"(one)); }   // IMAD.HI (1 instr)
            if (MODE == 4) { v[j] = lab_rotl_fma(v[j], pw); v[(j + 1) & 7] = lab_rotl(v[(j + 1) & 7], 7) ^ pw; }  // 2 FMA + 2 ALU
        }
    }
    uint32_t a = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) a ^= v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

// bare ChaCha20 double rounds on NB interleaved states per thread, nothing else: the practical ceiling of the round
// function's instruction mix (IMAD add : LOP3 xor : SHF rotate = 1 : 1 : 1) at a given number of resident warps
template <int NB, bool ALU_ADDS>
__global__ void __launch_bounds__(256) k_rounds(LabSeed seed, uint32_t *out, int double_rounds) {
    uint32_t x[NB][16];
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
        for (int i = 0; i < 16; i++) x[b][i] = threadIdx.x * 131u + blockIdx.x * 7u + b * 17u + i;
    if (ALU_ADDS) {
#pragma unroll 1
        for (int r = 0; r < double_rounds; r++) {
#pragma unroll
            for (int b = 0; b < NB; b++) {
#define QRA(a, b_, c, d) a += b_; d ^= a; d = lab_rotl(d, 16); c += d; b_ ^= c; b_ = lab_rotl(b_, 12); a += b_; d ^= a; d = lab_rotl(d, 8); c += d; b_ ^= c; b_ = lab_rotl(b_, 7);
                QRA(x[b][0], x[b][4], x[b][8], x[b][12]) QRA(x[b][1], x[b][5], x[b][9], x[b][13]) QRA(x[b][2], x[b][6], x[b][10], x[b][14]) QRA(x[b][3], x[b][7], x[b][11], x[b][15])
                QRA(x[b][0], x[b][5], x[b][10], x[b][15]) QRA(x[b][1], x[b][6], x[b][11], x[b][12]) QRA(x[b][2], x[b][7], x[b][8], x[b][13]) QRA(x[b][3], x[b][4], x[b][9], x[b][14])
            }
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < double_rounds; r++) lab_double_round<NB, 0u>(x, seed);
    }
    uint32_t a = 0;
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
        for (int i = 0; i < 16; i++) a ^= x[b][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    template <class F> float run(F f, int reps = 3) {
        f();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < reps; r++) {
            cudaEventRecord(a);
            f();
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        return best;
    }
};

static unsigned long long checksum(const uint32_t *d, size_t n) {
    unsigned long long *dc, h = 0;
    CK(cudaMalloc(&dc, 8));
    CK(cudaMemset(dc, 0, 8));
    k_checksum<<<592, 256>>>(d, n, dc);
    CK(cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost));
    cudaFree(dc);
    return h;
}

template <uint32_t RM>
static void bench_expand(Timer &tm, const LabSeed &seed, uint32_t *out, size_t n, uint64_t start, unsigned long long ref_small, size_t n_small) {
    float ms = tm.run([&] { k_crs_expand<RM><<<148 * 16, 256>>>(seed, start, 0ull, n, out); });
    CK(cudaGetLastError());
    k_crs_expand<RM><<<148 * 16, 256>>>(seed, start, 0ull, n_small, out);
    unsigned long long cs = checksum(out, n_small);
    printf("{\"kernel\": \"k_crs_expand\", \"rm\": \"0x%08x\", \"coeffs\": %zu, \"ms\": %.4f, \"blocks_per_s\": %.4e, \"exact\": %s}\n", RM, n, ms, n / (ms * 1e-3),
           cs == ref_small ? "true" : "false");
    fflush(stdout);
}

template <uint32_t RM, int PP>
static void bench_commit(Timer &tm, const LabSeed &seed, const uint32_t *What, uint32_t N, uint32_t R, uint64_t rows, uint32_t *T, unsigned long long *ref) {
    const unsigned grid = (unsigned)((rows + KA_RT - 1) / KA_RT);
    CK(cudaFuncSetAttribute(k_commit_inner<16, RM, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ka_dyn_smem(PP, 16)));
    auto launch = [&] {
        for (uint32_t ib = 0; ib < R; ib += KA_CONS * 16) k_commit_inner<16, RM, PP><<<grid, ka_threads(PP), ka_dyn_smem(PP, 16)>>>(seed, What, N, R, 0ull, rows, ib, T, rows, 0ull);
    };
    float ms = tm.run(launch, 2);
    CK(cudaGetLastError());
    unsigned long long cs = checksum(T, (size_t)R * rows * 64);
    if (!*ref) *ref = cs;
    printf("{\"kernel\": \"k_commit_inner\", \"pp\": %d, \"rm\": \"0x%08x\", \"N\": %u, \"R\": %u, \"rows\": %llu, \"ms\": %.4f, \"blocks_per_s\": %.4e, \"same_as_first\": %s}\n", PP, RM, N, R,
           (unsigned long long)rows, ms, (double)rows * N * 64 / (ms * 1e-3), cs == *ref ? "true" : "false");
    fflush(stdout);
}

int main(int argc, char **argv) {
    const LabSeed seed = mk_seed();
    Timer tm;
    const bool only_gen = argc > 1 && std::string(argv[1]) == "gen";      // `kbench gen`: only the limb-plane generator variants
    if (!only_gen) {   // pipe rates
        uint32_t *o;
        CK(cudaMalloc(&o, 148 * 16 * 256 * 4));
        const int iters = 4096;
        const double lane_ops = 148.0 * 8 * 256 * 8 * iters;
        const char *names[5] = {"SHF+LOP3 (2 ALU)", "IMAD (1 FMA)", "IMAD+IMAD.HI (2 FMA)", "IMAD.HI (1 FMA)", "2 FMA + 2 ALU interleaved"};
        const int instr[5] = {2, 1, 2, 1, 4};
        float ms[5];
        ms[0] = tm.run([&] { k_pipe<0><<<148 * 8, 256>>>(o, iters, 128u, 1u); });
        ms[1] = tm.run([&] { k_pipe<1><<<148 * 8, 256>>>(o, iters, 128u, 1u); });
        ms[2] = tm.run([&] { k_pipe<2><<<148 * 8, 256>>>(o, iters, 128u, 1u); });
        ms[3] = tm.run([&] { k_pipe<3><<<148 * 8, 256>>>(o, iters, 128u, 1u); });
        ms[4] = tm.run([&] { k_pipe<4><<<148 * 8, 256>>>(o, iters, 128u, 1u); });
        for (int m = 0; m < 5; m++)
            printf("{\"kernel\": \"k_pipe\", \"mode\": \"%s\", \"ms\": %.4f, \"lane_instr_per_s\": %.4e}\n", names[m], ms[m], lane_ops * instr[m] / (ms[m] * 1e-3));
        // round-function ceiling: blocks/s = threads * NB * (double_rounds / 10) / t
        const int dr = 2000;
        for (int bps = 1; bps <= 8; bps *= 2) {        // CTAs of 256 threads per SM = 2 * bps warps per scheduler
            const unsigned grid = 148 * bps;
            float t1 = tm.run([&] { k_rounds<1, false><<<grid, 256>>>(seed, o, dr); });
            float t2 = tm.run([&] { k_rounds<2, false><<<grid, 256>>>(seed, o, dr); });
            float t3 = tm.run([&] { k_rounds<3, false><<<grid, 256>>>(seed, o, dr); });
            float t4 = tm.run([&] { k_rounds<4, false><<<grid, 256>>>(seed, o, dr); });
            float t5 = tm.run([&] { k_rounds<2, true><<<grid, 256>>>(seed, o, dr); });
            const double blk = (double)grid * 256 * dr / 10.0;
            printf("{\"kernel\": \"k_rounds\", \"warps_per_scheduler\": %d, \"blocks_per_s_nb1\": %.4e, \"nb2\": %.4e, \"nb3\": %.4e, \"nb4\": %.4e, \"nb2_alu_adds\": %.4e}\n", 2 * bps,
                   blk / (t1 * 1e-3), 2 * blk / (t2 * 1e-3), 3 * blk / (t3 * 1e-3), 4 * blk / (t4 * 1e-3), 2 * blk / (t5 * 1e-3));
            fflush(stdout);
        }
        cudaFree(o);
    }
    if (!only_gen) {   // expansion
        const size_t n = (size_t)1 << 26, n_small = (size_t)1 << 20;
        // start chosen so that low32(seed + counter) crosses a 2^32 boundary inside the small range
        const uint64_t low = (uint32_t)seed.limb[0];
        const uint64_t start = (0x100000000ull - low) - 1000ull + 0x500000000ull;
        uint32_t *out;
        CK(cudaMalloc(&out, n * 4));
        k_expand_generic<<<(unsigned)((n_small + 255) / 256), 256>>>(seed, start, n_small, out);
        const unsigned long long ref = checksum(out, n_small);
        bench_expand<0x00000000u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x00010001u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x01010101u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x01110111u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x11111111u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x11151115u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x15151515u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x55555555u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x44444444u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x22222222u>(tm, seed, out, n, start, ref, n_small);
        bench_expand<0x88888888u>(tm, seed, out, n, start, ref, n_small);
        cudaFree(out);
    }
    {   // generation into limb planes: occupancy variants
        const uint32_t N = 4096;
        const uint64_t rows = 148 * 32;
        const uint32_t kpad = 2 * N, nt = (uint32_t)(rows / 64);
        uint8_t *planes;
        CK(cudaMalloc(&planes, (size_t)nt * 64 * 32 * 2 * kpad));
        auto report = [&](const char *name, float ms) {
            printf("{\"kernel\": \"k_gen_planes\", \"variant\": \"%s\", \"rows\": %llu, \"N\": %u, \"ms\": %.4f, \"blocks_per_s\": %.4e}\n", name, (unsigned long long)rows, N, ms,
                   (double)rows * N * 64 / (ms * 1e-3));
            fflush(stdout);
        };
        {
            int nb2 = 0, nb3 = 0, nb4 = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_gen_planes<2>, 32 * GP_WARPS, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb3, k_gen_planes<3>, 32 * GP_WARPS, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb4, k_gen_planes<4>, 32 * GP_WARPS, 0);
            printf("{\"kernel\": \"k_gen_planes\", \"resident_ctas_per_sm\": {\"minblocks2\": %d, \"minblocks3\": %d, \"minblocks4\": %d}}\n", nb2, nb3, nb4);
        }
        const size_t plane_words = (size_t)nt * 64 * 32 * 2 * kpad / 4;
        auto sum = [&]() { return checksum((const uint32_t *)planes, plane_words); };
        report("reference for the checksums: var 0, minblocks 3, grid 148 x 12", tm.run([&] { k_gen_planes<3, 0><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); }, 2));
        const unsigned long long ref = sum();
        auto variant = [&](const char *name, auto launch) {
            CK(cudaMemset(planes, 0, plane_words * 4));
            const float ms = tm.run(launch, 2);
            const bool same = sum() == ref;
            printf("{\"kernel\": \"k_gen_planes\", \"variant\": \"%s\", \"ms\": %.4f, \"blocks_per_s\": %.4e, \"same_planes\": %s}\n", name, ms,
                   (double)rows * N * 64 / (ms * 1e-3), same ? "true" : "false");
            fflush(stdout);
        };
        variant("library launch (LAB_GP_VAR, LAB_GP_MINB, 8 waves)", [&] { k_gen_planes<LAB_GP_MINB><<<148 * LAB_GP_MINB * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 0 (packed shuffle, 64-bit tag check per polynomial)", [&] { k_gen_planes<3, 0><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 1 (split shuffles)", [&] { k_gen_planes<3, 1><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 2 (32-bit window check inside a run)", [&] { k_gen_planes<3, 2><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 3 (both)", [&] { k_gen_planes<3, 3><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 5 (split shuffles, double rounds unrolled x2)", [&] { k_gen_planes<3, 5><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 9 (split shuffles, double rounds unrolled x4)", [&] { k_gen_planes<3, 9><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 13 (split shuffles, double rounds unrolled x8)", [&] { k_gen_planes<3, 13><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 29 (two polynomials per warp iteration: 4 ChaCha20 states per lane), minblocks 2", [&] { k_gen_planes<2, 29><<<148 * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 29, minblocks 3", [&] { k_gen_planes<3, 29><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 21 (x2, unroll x2), minblocks 2", [&] { k_gen_planes<2, 21><<<148 * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 25 (x2, unroll x4), minblocks 2", [&] { k_gen_planes<2, 25><<<148 * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 21 (x2, unroll x2), minblocks 3", [&] { k_gen_planes<3, 21><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 25 (x2, unroll x4), minblocks 3", [&] { k_gen_planes<3, 25><<<148 * 12, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 21 (x2, unroll x2), minblocks 2, grid 148 x 2", [&] { k_gen_planes<2, 21><<<148 * 2, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 21 (x2, unroll x2), minblocks 2, grid 148 x 16", [&] { k_gen_planes<2, 21><<<148 * 16, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 17 (x2, rolled), minblocks 2", [&] { k_gen_planes<2, 17><<<148 * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 0, minblocks 4 (32 warps/SM)", [&] { k_gen_planes<4, 0><<<148 * 16, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 3, minblocks 4 (32 warps/SM)", [&] { k_gen_planes<4, 3><<<148 * 16, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 0, minblocks 2 (16 warps/SM)", [&] { k_gen_planes<2, 0><<<148 * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        variant("var 3, minblocks 2 (16 warps/SM)", [&] { k_gen_planes<2, 3><<<148 * 8, 32 * GP_WARPS>>>(seed, N, 0ull, rows, planes, nt, kpad); });
        CK(cudaGetLastError());
        cudaFree(planes);
    }
    if (!only_gen) {   // inner commitment
        const uint32_t N = argc > 1 ? atoi(argv[1]) : 1024, R = 64;
        const uint64_t rows = 148 * 4 * 4;
        const size_t hats = (size_t)(N + KA_PAD_COLS) * R + KA_PAD_VECS;
        uint32_t *What, *T;
        CK(cudaMalloc(&What, hats * 128));
        CK(cudaMalloc(&T, (size_t)R * rows * 256));
        k_fill_hat<<<(unsigned)((hats * 32 + 255) / 256), 256>>>(What, hats * 32);
        unsigned long long ref = 0;
        bench_commit<0x00000000u, 1>(tm, seed, What, N, R, rows, T, &ref);
        bench_commit<0x00000000u, 2>(tm, seed, What, N, R, rows, T, &ref);
        bench_commit<0x00010001u, 2>(tm, seed, What, N, R, rows, T, &ref);
        cudaFree(What); cudaFree(T);
    }
    return 0;
}
