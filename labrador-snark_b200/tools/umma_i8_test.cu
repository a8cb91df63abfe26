// Bring-up test of the tcgen05 int8 path (developer tool): D[128 x 256] (s32, TMEM) = A[128 x K] * B[256 x K]^T with
// K-major int8 operands in 128-byte-swizzled shared memory, one CTA, operands placed by plain stores (no TMA here).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o bin/umma_i8_test umma_i8_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int M = 128, N = 256, KC = 128;       // KC = K elements per smem tile row (128 bytes = one swizzle row)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups of 1024 bytes (stride byte offset), 16-byte chunks XOR-ed with row % 8
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128) k_umma(const int8_t *A, const int8_t *B, int32_t *D, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                 // [M][128]
    uint8_t *sB = smem + M * 128;       // [N][128]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tmem_base;
    // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 256, M = 128
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
        // stage the K-chunk: element (r, c) at r * 128 + ((c / 16) ^ (r % 8)) * 16 + c % 16
        for (int idx = tid; idx < (M + N) * 8; idx += 128) {
            const int r = idx >> 3, ch = idx & 7;
            const bool isA = r < M;
            const int rr = isA ? r : r - M;
            const int8_t *src = (isA ? A : B) + (size_t)rr * K + k0 + ch * 16;
            uint8_t *dst = (isA ? sA : sB) + rr * 128 + ((ch ^ (rr & 7)) * 16);
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src);
        }
        asm volatile("fence.proxy.async.shared::cta;");        // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
#pragma unroll
            for (int k = 0; k < KC / 32; k++) {
                const uint32_t acc = (k0 > 0 || k > 0) ? 1u : 0u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tm),
                    "l"(da + (uint64_t)(k * 32 / 16)), "l"(db + (uint64_t)(k * 32 / 16)), "r"(idesc), "r"(acc));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bar)));
        }
        // everyone waits until the MMAs have consumed the smem tiles
        asm volatile(
            "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra W;\nDONE:\n}" ::"r"(smem_u32(&bar)), "r"(phase));
        phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= rows), 32 columns per load
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
              "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
              "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        const int row = warp * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; j++) D[(size_t)row * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256));
}

int main() {
    const int K = 512;
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)(s >> 24); };
    for (auto &x : hA) x = (int8_t)(rnd() % 128);                 // limbs of A: 0..127
    for (auto &x : hB) x = (int8_t)(rnd() % 255 - 127);           // signed limbs of S: -127..127
    int8_t *dA, *dB;
    int32_t *dD;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
    const int smem = (M + N) * 128 + 1024;
    CK(cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_umma<<<1, 128, smem>>>(dA, dB, dD, K);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> hD((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            int32_t ref = 0;
            for (int k = 0; k < K; k++) ref += (int32_t)hA[(size_t)m * K + k] * (int32_t)hB[(size_t)n * K + k];
            if (ref != hD[(size_t)m * N + n]) {
                if (bad < 8) printf("mismatch at (%d,%d): got %d want %d\n", m, n, hD[(size_t)m * N + n], ref);
                bad++;
            }
        }
    printf("{\"test\": \"umma_i8 128x256x%d\", \"mismatches\": %ld}\n", K, bad);
    return bad != 0;
}
