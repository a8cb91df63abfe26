// sm_100a kernels of the LaBRADOR prover hot path.  All arithmetic is exact integer arithmetic.
//
// Device layouts
//   poly  : uint32_t[64] canonical coefficients                              (API boundary)
//   hat   : uint32_t[32] packed slots re | im << 16, canonical              (transform domain)
//   What  : witness in the transform domain, n-major: hat[(n * R + i)]       (K_A reads 128 B / (n,i))
//
// Kernels (roofline that bounds each is named; see DESIGN.md for the byte/op counts)
//   k_ntt_fwd_regs / k_ntt_inv_regs / k_polymul_regs   lane-per-poly, swizzled smem staging   HBM
//   k_crs_expand                                       thread per 2 coefficients (trimmed ChaCha20) INT32 ALU
//   k_commit_inner (K_A)                               warp-specialised: ChaCha20 + warp NTT producers,
//                                                      TMA-fed IMAD consumers; shapes < 2^22 polynomials  INT32 ALU
//   k_crs_matvec<FILL_CACHE> + k_finish_rows (K_MV)    CRS gen + warp NTT + mat-vec (optionally write-through) INT32 ALU
//   k_cached_matvec                                    the same mat-vec from the CRS cache     HBM
//   k_fwd_hat / k_inv_hat / k_decomp_fwd / k_ip_hat / k_pointwise / ...   HBM / L2  (k_jl2, k_piT_omega2: lab_jl.cuh)
// lab_umma.cuh: k_gen_planes (ChaCha20 -> int8 limb planes, INT32 ALU) + k_umma_commit (tcgen05 contraction, HBM) --
// the large-shape inner commitment; lab_gen.cuh: device-side challenge / witness / statement generation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lab_chacha.cuh"
#include "lab_ntt.cuh"
#include "lab_crsgen.cuh"

namespace lab {

// ------------------------------------------------------------------------------------------------
// lane-per-poly batch transforms.  A CTA of 128 threads stages 128 polys (32 KB) through shared
// memory with a 16-byte-chunk XOR swizzle so that both the coalesced global side (consecutive
// lanes -> consecutive chunks) and the per-lane side (lane p reads chunk c of its own poly) are
// bank-conflict free:  chunk (p, c) lives at p * 16 + (c ^ (p & 7))   [c in 0..15, low 3 bits swizzled]
// ------------------------------------------------------------------------------------------------
constexpr int NTT_TPB = 128;

__device__ __forceinline__ int swz(int p, int c) { return p * 16 + (c ^ (p & 7)); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// asynchronous variant of stage_in: global -> shared without passing through registers, so the issuing thread can go on
// computing; complete with cp_async_wait_all() + __syncthreads().  Out-of-range chunks are zero-filled synchronously.
__device__ __forceinline__ void stage_in_async(uint4 *sm, const uint32_t *__restrict__ g, size_t base_poly, size_t n_polys, int tid) {
    const uint4 *g4 = reinterpret_cast<const uint4 *>(g) + base_poly * 16;
    const size_t avail = (n_polys - base_poly) * 16;
#pragma unroll
    for (int it = 0; it < 16; it++) {
        int q = it * 128 + tid;
        uint4 *dst = &sm[swz(q >> 4, q & 15)];
        if ((size_t)q < avail) cp_async16(dst, g4 + q);
        else *dst = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
}

__device__ __forceinline__ void stage_in(uint4 *sm, const uint32_t *__restrict__ g, size_t base_poly, size_t n_polys, int tid) {
    // 128 polys * 16 chunks = 2048 chunks, 16 per thread, coalesced
    const uint4 *g4 = reinterpret_cast<const uint4 *>(g) + base_poly * 16;
    const size_t avail = (n_polys - base_poly) * 16;
#pragma unroll
    for (int it = 0; it < 16; it++) {
        int q = it * NTT_TPB + tid;
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((size_t)q < avail) v = __ldg(g4 + q);
        sm[swz(q >> 4, q & 15)] = v;
    }
}
__device__ __forceinline__ void stage_out(const uint4 *sm, uint32_t *__restrict__ g, size_t base_poly, size_t n_polys, int tid) {
    uint4 *g4 = reinterpret_cast<uint4 *>(g) + base_poly * 16;
    const size_t avail = (n_polys - base_poly) * 16;
#pragma unroll
    for (int it = 0; it < 16; it++) {
        int q = it * NTT_TPB + tid;
        if ((size_t)q < avail) g4[q] = sm[swz(q >> 4, q & 15)];
    }
}
// coefficient order -> registers: re[d] = f_d, im[d] = f_{d+32}
__device__ __forceinline__ void load_coeffs(const uint4 *sm, int p, uint32_t (&re)[32], uint32_t (&im)[32]) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint4 a = sm[swz(p, c)], b = sm[swz(p, c + 8)];
        re[4 * c] = a.x; re[4 * c + 1] = a.y; re[4 * c + 2] = a.z; re[4 * c + 3] = a.w;
        im[4 * c] = b.x; im[4 * c + 1] = b.y; im[4 * c + 2] = b.z; im[4 * c + 3] = b.w;
    }
}
__device__ __forceinline__ void store_coeffs(uint4 *sm, int p, const uint32_t (&re)[32], const uint32_t (&im)[32]) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
        sm[swz(p, c)] = make_uint4(re[4 * c], re[4 * c + 1], re[4 * c + 2], re[4 * c + 3]);
        sm[swz(p, c + 8)] = make_uint4(im[4 * c], im[4 * c + 1], im[4 * c + 2], im[4 * c + 3]);
    }
}
// slot order, interleaved (re_j, im_j) -> registers
__device__ __forceinline__ void load_slots(const uint4 *sm, int p, uint32_t (&re)[32], uint32_t (&im)[32]) {
#pragma unroll
    for (int c = 0; c < 16; c++) {
        uint4 a = sm[swz(p, c)];
        re[2 * c] = a.x; im[2 * c] = a.y; re[2 * c + 1] = a.z; im[2 * c + 1] = a.w;
    }
}
__device__ __forceinline__ void store_slots(uint4 *sm, int p, const uint32_t (&re)[32], const uint32_t (&im)[32]) {
#pragma unroll
    for (int c = 0; c < 16; c++) sm[swz(p, c)] = make_uint4(re[2 * c], im[2 * c], re[2 * c + 1], im[2 * c + 1]);
}
__device__ __forceinline__ void canon_in(uint32_t (&re)[32], uint32_t (&im)[32]) {
    // API inputs are documented canonical (< Q).  The generated transforms are bound-checked for 14-bit inputs, so
    // canonical data passes untouched; if any value is larger (caller error, but still well defined: the value mod Q
    // is what counts) this lane reduces its polynomial first.  The branch is uniform on valid data.
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) any |= re[j] | im[j];
    if (any >> 14) {
#pragma unroll
        for (int j = 0; j < 32; j++) { re[j] = lab_fold(lab_fold(re[j])); im[j] = lab_fold(lab_fold(im[j])); }
    }
}

__global__ void __launch_bounds__(NTT_TPB) k_ntt_fwd_regs(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_polys) {
    __shared__ uint4 sm[NTT_TPB * 16];
    const int tid = threadIdx.x;
    for (size_t base = (size_t)blockIdx.x * NTT_TPB; base < n_polys; base += (size_t)gridDim.x * NTT_TPB) {
        stage_in(sm, in, base, n_polys, tid);
        __syncthreads();
        uint32_t re[32], im[32];
        load_coeffs(sm, tid, re, im);
        canon_in(re, im);
        lab_ntt32_fwd_regs(re, im);
        __syncthreads();
        store_slots(sm, tid, re, im);
        __syncthreads();
        stage_out(sm, out, base, n_polys, tid);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(NTT_TPB) k_ntt_inv_regs(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_polys) {
    __shared__ uint4 sm[NTT_TPB * 16];
    const int tid = threadIdx.x;
    for (size_t base = (size_t)blockIdx.x * NTT_TPB; base < n_polys; base += (size_t)gridDim.x * NTT_TPB) {
        stage_in(sm, in, base, n_polys, tid);
        __syncthreads();
        uint32_t re[32], im[32];
        load_slots(sm, tid, re, im);
        canon_in(re, im);
        lab_ntt32_inv_regs(re, im);
        __syncthreads();
        store_coeffs(sm, tid, re, im);
        __syncthreads();
        stage_out(sm, out, base, n_polys, tid);
        __syncthreads();
    }
}
// c = a * b in R_q: two forward transforms, 32 slot products, one inverse -- &Rq * &Rq (algebraic.rs:517-523).
// The transformed first operand waits in shared memory (packed, [slot][thread]: conflict free) while the second one is
// transformed, which keeps the kernel at 4 CTAs per SM without spills.
__global__ void __launch_bounds__(NTT_TPB, 4) k_polymul_regs(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                              uint32_t *__restrict__ c, size_t n_polys) {
    __shared__ uint4 sm[NTT_TPB * 16];
    __shared__ uint32_t ahat[32][NTT_TPB];
    const int tid = threadIdx.x;
    for (size_t base = (size_t)blockIdx.x * NTT_TPB; base < n_polys; base += (size_t)gridDim.x * NTT_TPB) {
        uint32_t re[32], im[32];
        stage_in_async(sm, a, base, n_polys, tid);
        cp_async_wait_all();
        __syncthreads();
        load_coeffs(sm, tid, re, im);
        __syncthreads();
        stage_in_async(sm, b, base, n_polys, tid);   // lands while the first operand is transformed
        canon_in(re, im);
        lab_ntt32_fwd_regs_lazy(re, im);             // < 2Q is enough for the slot products
#pragma unroll
        for (int j = 0; j < 32; j++) ahat[j][tid] = lab_pack(re[j], im[j]);
        cp_async_wait_all();
        __syncthreads();
        load_coeffs(sm, tid, re, im);
        canon_in(re, im);
        lab_ntt32_fwd_regs_lazy(re, im);
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const uint32_t pk = ahat[j][tid];
            uint32_t r, i;
            lab_cmul(lab_re(pk), lab_im(pk), re[j], im[j], r, i);
            re[j] = r; im[j] = i;                // < 2Q, inside the generated inverse's input bound
        }
        lab_ntt32_inv_regs(re, im);
        __syncthreads();
        store_coeffs(sm, tid, re, im);
        __syncthreads();
        stage_out(sm, c, base, n_polys, tid);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// warp-per-poly helpers (lane j <-> coefficient pair (j, j+32) / slot j)
// ------------------------------------------------------------------------------------------------
// poly (canonical u32[64]) -> hat (packed u32[32]); out index = perm(p) to allow the n-major transpose:
// p = o * inner + i  ->  i * outer + o   (inner == 0: identity)
__global__ void __launch_bounds__(256) k_fwd_hat(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_polys,
                                                 size_t inner, size_t outer) {
    const int lane = threadIdx.x & 31;
    const LabWarpTw tw = lab_warp_tw(lane);
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t p = warp; p < n_polys; p += nw) {
        uint32_t re = lab_fold(lab_fold(in[p * 64 + lane])), im = lab_fold(lab_fold(in[p * 64 + 32 + lane]));
        lab_ntt32_fwd_warp(re, im, tw, lane);
        size_t q = inner ? (p % inner) * outer + (p / inner) : p;
        out[q * 32 + lane] = lab_pack(re, im);
    }
}
__global__ void __launch_bounds__(256) k_inv_hat(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_polys) {
    const int lane = threadIdx.x & 31;
    const LabWarpTw tw = lab_warp_tw(lane);
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t p = warp; p < n_polys; p += nw) {
        uint32_t v = in[p * 32 + lane];
        uint32_t re = lab_re(v), im = lab_im(v);
        lab_ntt32_inv_warp(re, im, tw, lane);
        out[p * 64 + lane] = re;
        out[p * 64 + 32 + lane] = im;
    }
}

// digit_k(c) = f(floor(c / b^k) mod b), f(x) = x if x <= floor(b/2) else b - x  (util.rs:360-442)
__device__ __forceinline__ uint32_t digit_of(uint32_t &v, uint32_t base) {
    uint32_t q = v / base, x = v - q * base;
    v = q;
    return x <= (base >> 1) ? x : base - x;
}
// decompose + forward transform.  in: poly[n_polys]; out: hat[((p / inner) * exp + k) * inner + (p % inner)]
__global__ void __launch_bounds__(256) k_decomp_fwd(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_polys,
                                                    size_t inner, uint32_t base, int exp) {
    const int lane = threadIdx.x & 31;
    const LabWarpTw tw = lab_warp_tw(lane);
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t p = warp; p < n_polys; p += nw) {
        uint32_t a = in[p * 64 + lane], b = in[p * 64 + 32 + lane];
        for (int k = 0; k < exp; k++) {
            uint32_t re = digit_of(a, base), im = digit_of(b, base);
            re = lab_canon(re); im = lab_canon(im);     // digits can exceed Q only for base > 2Q
            lab_ntt32_fwd_warp(re, im, tw, lane);
            size_t q = ((p / inner) * (size_t)exp + (size_t)k) * inner + (p % inner);
            out[q * 32 + lane] = lab_pack(re, im);
        }
    }
}
// plain coefficient-domain decomposition (lab_decompose): out[k][p][d]
__global__ void k_decompose(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_coeffs, uint32_t base, int exp) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n_coeffs; idx += stride) {
        uint32_t v = in[idx];
        for (int k = 0; k < exp; k++) out[(size_t)k * n_coeffs + idx] = lab_canon(digit_of(v, base));
    }
}
__global__ void k_sigma_inv(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n_coeffs) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n_coeffs; idx += stride) {
        size_t p = idx >> 6;
        int d = (int)(idx & 63);
        uint32_t v = lab_canon(in[p * 64 + ((64 - d) & 63)]);
        out[idx] = d == 0 ? v : (v ? LABQ - v : 0u);
    }
}

// ------------------------------------------------------------------------------------------------
// batched inner products in the transform domain:
//   out[b][slot] = scale * sum_{n < len} X[(n * x_sn + xi(b) * x_si)] (*) Y[(n * y_sn + yi(b) * y_si)]
// with b = xi * nby + yi.  One CTA per b, 8 warps stride over n, smem tree at the end.
// mode 0: plain; mode 2: diagonal (xi = yi = b); mode 1: symmetrised h: out[b] = (X[xi].Y[yi] + X[yi].Y[xi]) * scale  (proofgen.rs:334-347)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ip_hat(const uint32_t *__restrict__ X, size_t x_sn, size_t x_si,
                                                const uint32_t *__restrict__ Y, size_t y_sn, size_t y_si,
                                                size_t len, size_t nby, uint32_t scale, int mode, uint32_t *__restrict__ out, size_t xi0 = 0) {
    __shared__ uint32_t sre[8][32], sim[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t b = blockIdx.x;
    const size_t xi = mode == 2 ? b : b / nby + xi0, yi = mode == 2 ? b : b % nby;       // xi0: first row of a row shard (out stays local)
    uint32_t accr = 0, acci = 0;
    int pending = 0;
    for (size_t n = w; n < len; n += 8) {
        uint32_t x = X[(n * x_sn + xi * x_si) * 32 + lane], y = Y[(n * y_sn + yi * y_si) * 32 + lane];
        uint32_t xr = lab_re(x), xm = lab_im(x), yr = lab_re(y), ym = lab_im(y);
        accr += xr * yr + (LABQ - xm) * ym;
        acci += xr * ym + xm * yr;
        if (mode == 1) {
            uint32_t x2 = X[(n * x_sn + yi * x_si) * 32 + lane], y2 = Y[(n * y_sn + xi * y_si) * 32 + lane];
            uint32_t ar = lab_re(x2), am = lab_im(x2), br = lab_re(y2), bm = lab_im(y2);
            accr += ar * br + (LABQ - am) * bm;
            acci += ar * bm + am * br;
        }
        if (++pending == 8) { accr = lab_fold(accr); acci = lab_fold(acci); pending = 0; }   // 8 * 4 * 2^26 < 2^32
    }
    sre[w][lane] = lab_canon(accr);
    sim[w][lane] = lab_canon(acci);
    __syncthreads();
    if (w == 0) {
        uint32_t r = 0, i = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { r += sre[k][lane]; i += sim[k][lane]; }
        r = lab_canon(lab_canon(r) * scale);
        i = lab_canon(lab_canon(i) * scale);
        out[b * 32 + lane] = lab_pack(r, i);
    }
}

// generic slot-wise fused multiply-add over hats:
//   out[p] = A[ai(p)] (*) X[p]  (+ B[bi(p)] (*) Y[p] if Y)      ai(p) = (p / a_div) % a_mod  (a_mod == 0: index 0)
__global__ void k_pointwise(const uint32_t *__restrict__ A, size_t a_div, size_t a_mod, const uint32_t *__restrict__ X,
                            const uint32_t *__restrict__ B, size_t b_div, size_t b_mod, const uint32_t *__restrict__ Y,
                            uint32_t *__restrict__ out, size_t n_polys) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n_polys * 32; idx += stride) {
        size_t p = idx >> 5;
        int s = (int)(idx & 31);
        size_t ai = a_mod ? (p / a_div) % a_mod : 0;
        uint32_t a = A[ai * 32 + s], x = X[idx];
        uint32_t r, i;
        lab_cmul(lab_re(a), lab_im(a), lab_re(x), lab_im(x), r, i);
        if (Y) {
            size_t bi = b_mod ? (p / b_div) % b_mod : 0;
            uint32_t b = B[bi * 32 + s], y = Y[idx];
            uint32_t r2, i2;
            lab_cmul(lab_re(b), lab_im(b), lab_re(y), lab_im(y), r2, i2);
            r += r2; i += i2;
        }
        out[idx] = lab_pack(lab_canon(r), lab_canon(i));
    }
}
// out[n] = sum_{i in [i0,i0+ni)} C[i] (*) W[n * R + i]   (amortised opening z, proofgen.rs:387-399)
__global__ void __launch_bounds__(256) k_amortize(const uint32_t *__restrict__ Chat, const uint32_t *__restrict__ What, size_t N, size_t R,
                                                  size_t i0, size_t ni, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t n = warp; n < N; n += nw) {
        uint32_t accr = 0, acci = 0;
        int pending = 0;
        for (size_t i = i0; i < i0 + ni; i++) {
            uint32_t c = Chat[i * 32 + lane], s = What[(n * R + i) * 32 + lane];
            uint32_t cr = lab_re(c), cm = lab_im(c), sr = lab_re(s), sm_ = lab_im(s);
            accr += cr * sr + (LABQ - cm) * sm_;
            acci += cr * sm_ + cm * sr;
            if (++pending == 16) { accr = lab_fold(accr); acci = lab_fold(acci); pending = 0; }
        }
        out[n * 32 + lane] = lab_pack(lab_canon(accr), lab_canon(acci));
    }
}
// sum over `cnt` hats spaced `stride` apart: out[p] = sum_t in[p + t * stride]   (b'' accumulation etc.)
__global__ void k_sum_hats(const uint32_t *__restrict__ in, size_t cnt, size_t stride_polys, uint32_t *__restrict__ out, size_t n_polys) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_polys * 32) return;
    uint32_t r = 0, i = 0;
    for (size_t t = 0; t < cnt; t++) {
        uint32_t v = in[idx + t * stride_polys * 32];
        r += lab_re(v); i += lab_im(v);
        if ((t & 1023) == 1023) { r = lab_fold(r); i = lab_fold(i); }
    }
    out[idx] = lab_pack(lab_canon(r), lab_canon(i));
}

// number of differing words between two buffers (verifier equality checks 15-20, verification.rs:291-435)
__global__ void k_count_diff(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t n, unsigned long long *out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long d = 0;
    for (; idx < n; idx += stride) d += a[idx] != b[idx];
#pragma unroll
    for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0 && d) atomicAdd(out, d);
}
// out[p] = X[p] + Y[p] - Z[p] (mod Q) over hats; Y or Z may be null
__global__ void k_addsub_hats(const uint32_t *__restrict__ X, const uint32_t *__restrict__ Y, const uint32_t *__restrict__ Z, uint32_t *__restrict__ out, size_t n_words) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_words) return;
    uint32_t x = X[idx], r = lab_re(x), i = lab_im(x);
    if (Y) { uint32_t y = Y[idx]; r += lab_re(y); i += lab_im(y); }
    if (Z) { uint32_t z = Z[idx]; r += 2u * LABQ - lab_re(z); i += 2u * LABQ - lab_im(z); }
    out[idx] = lab_pack(lab_canon(r), lab_canon(i));
}

// c = a + b or a - b, coefficientwise mod Q (algebraic.rs:441-515)
__global__ void k_rq_addsub(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t *__restrict__ c, size_t n, int sub) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n; idx += stride) {
        const uint32_t x = lab_canon(a[idx]), y = lab_canon(b[idx]);
        c[idx] = lab_csub(sub ? x + LABQ - y : x + y);
    }
}
// exact sum of squares of canonical representatives (util.rs:195-202)
__global__ void __launch_bounds__(256) k_norm_sq(const uint32_t *__restrict__ in, size_t n, unsigned long long *out) {
    unsigned long long acc = 0;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n; idx += stride) {
        unsigned long long v = in[idx];
        acc += v * v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
// sum of squared digits of every coefficient (verifier Check 14, verification.rs:185-267), exact
__global__ void __launch_bounds__(256) k_digit_norm_sq(const uint32_t *__restrict__ in, size_t n, uint32_t base, int exp, unsigned long long *out) {
    unsigned long long acc = 0;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n; idx += stride) {
        uint32_t v = in[idx];
        for (int k = 0; k < exp; k++) {
            unsigned long long dg = lab_canon(digit_of(v, base));
            acc += dg * dg;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------------
// CRS expansion (structs.rs:35-45,147-171): thread per coefficient, coalesced stores.  INT32-ALU bound.
// ------------------------------------------------------------------------------------------------
// FMA-pipe rotation masks (lab_chacha.cuh), one per kernel.  Measured on a B200 (tools/kbench.cu,
// profiles/kbench_r1_rotation_masks.jsonl): IMAD.HI is half rate, so a rotation on the FMA pipe costs three issue slots'
// worth of pipe time -- it gains 3 % in k_crs_expand with two of the 32 rotations of a double round moved and loses
// everywhere else, hence 0 for the kernels that also multiply.
#ifndef LAB_RM_EXPAND
#define LAB_RM_EXPAND 0x00000000u
#endif
#ifndef LAB_RM_COMMIT
#define LAB_RM_COMMIT 0x00000000u
#endif
#ifndef LAB_RM_MATVEC
#define LAB_RM_MATVEC 0x00000000u
#endif

// LAB_KA_VAR (K_A producers and k_crs_expand): bit 0 = split shuffles in the transform, bits 2-3 = log2 of the unroll factor of the
// double-round loop (see LAB_MV_VAR below and LAB_GP_VAR in lab_umma.cuh)
#ifndef LAB_KA_VAR
#define LAB_KA_VAR 0         /* measured (profiles/kbench_r2b_ka_variants.jsonl): 13 is 3 % slower in K_A (80-register producers) and no faster in k_crs_expand */
#endif
template <uint32_t RM>
__global__ void __launch_bounds__(256) k_crs_expand(LabSeed seed, uint64_t start_lo, uint64_t start_hi, size_t n_coeffs, uint32_t *__restrict__ out) {
    size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
    LabHoist h;
    lab_hoist_invalidate(h);
    for (; idx < n_coeffs; idx += stride) {
        uint64_t lo = start_lo + idx;
        uint64_t hi = start_hi + (lo < start_lo);
        const uint32_t off[2] = {0u, 1u};
        uint32_t c[2];
        lab_crs_coeffs<2, RM, 1 << ((LAB_KA_VAR >> 2) & 3)>(seed, h, lo, hi, off, c);
        if (idx + 1 < n_coeffs) *reinterpret_cast<uint2 *>(out + idx) = make_uint2(c[0], c[1]);
        else out[idx] = c[0];
    }
}

// one CRS polynomial per warp, in the transform domain: lane j produces coefficients j and j+32
// LAB_MV_VAR (K_MV and the hat generator of the proof graphs): bit 0 = split shuffles in the transform, bits 2-3 = log2 of the
// unroll factor of the double-round loop
#ifndef LAB_MV_VAR
#define LAB_MV_VAR 13        /* measured on cfg 5 (1024 default statements): 0 -> 2984, 1 -> 3005, 5 -> 3043, 13 -> 3122 proofs/s */
#endif
template <uint32_t RM>
__device__ __forceinline__ void crs_poly_hat(const LabSeed &seed, LabHoist &h, uint64_t lo, uint64_t hi, const LabWarpTw &tw, int lane, uint32_t &re, uint32_t &im) {
    const uint32_t off[2] = {(uint32_t)lane, (uint32_t)lane + 32u};
    uint32_t c[2];
    lab_crs_coeffs<2, RM, 1 << ((LAB_MV_VAR >> 2) & 3)>(seed, h, lo, hi, off, c);
    re = c[0]; im = c[1];
    lab_ntt32_fwd_warp<(LAB_MV_VAR & 1) != 0>(re, im, tw, lane, seed.one);
}

// ------------------------------------------------------------------------------------------------
// K_A: inner Ajtai commitments t_i[row] = sum_n A[row][n] * s_i[n]   (proofgen.rs:41-49)
// One CTA owns KA_RT = 4 rows of A and walks all N columns in tiles of `cols` columns.  A is generated exactly once.
//
// Warp-specialised: the "producer" warps (registers trimmed with setmaxnreg) do nothing but ChaCha20 + the warp
// transform, PP polynomials of A per tile each, and park them (re, im, Q - im) in a 4-deep shared-memory ring; the four
// "consumer" warps (one warpgroup with a raised register budget) hold the accumulators (4 rows x IC witness vectors per
// thread, lane = slot) and multiply-accumulate every tile against the n-major transformed witness (128 B coalesced per
// (n, i), L2 resident), fetched one column ahead.  Producers never wait for the MACs (ring + FULL named barriers /
// EMPTY mbarriers), so the ALU pipe sees an uninterrupted ChaCha20 stream while the consumers' IMADs use the FMA pipe.
//   PP = 1: 12 producer warps x 1 polynomial (2 interleaved ChaCha20 states per lane), tile = 4 rows x 3 columns
//   PP = 2:  8 producer warps x 2 polynomials (4 interleaved states per lane),          tile = 4 rows x 4 columns
// ------------------------------------------------------------------------------------------------
constexpr int KA_RT = 4;        // rows of A per CTA
constexpr int KA_CONS = 4;      // consumer warps (one warpgroup)
constexpr int KA_DEPTH = 4;     // ring slots between producer and consumer warps
__host__ __device__ constexpr int ka_prod(int pp) { return pp == 1 ? 12 : 8; }            // producer warps
__host__ __device__ constexpr int ka_cols(int pp) { return ka_prod(pp) * pp / KA_RT; }    // columns of A per tile
__host__ __device__ constexpr int ka_threads(int pp) { return 32 * (ka_prod(pp) + KA_CONS); }
__host__ __device__ constexpr size_t ka_dyn_smem(int pp, int ic) { return (size_t)KA_DEPTH * ka_cols(pp) * KA_CONS * ic * 128; }   // witness ring
#ifndef LAB_KA_PP
#define LAB_KA_PP 1
#endif
// hats of padding the transformed witness needs after its N*R entries: the consumers read up to `cols` columns past
// the end (ring tail + one-column prefetch) and up to KA_CONS*16 vectors past R without predicates
constexpr size_t KA_PAD_COLS = 4 + 1;
constexpr size_t KA_PAD_VECS = KA_CONS * 16;

__device__ __forceinline__ void bar_sync_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (cp.async.bulk, sm_90+)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
// 16-bit shared load zero-extended into a 32-bit register (no ALU-pipe unpacking)
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

template <int IC, uint32_t RM, int PP>
__global__ void __launch_bounds__(ka_threads(PP), 1) k_commit_inner(LabSeed seed, const uint32_t *__restrict__ What, uint32_t N, uint32_t R,
                                                                     uint64_t row0, uint64_t nrows, uint32_t i_base, uint32_t *__restrict__ T,
                                                                     uint64_t t_stride, uint64_t t_row_off) {
    constexpr int PROD = ka_prod(PP), COLS = ka_cols(PP), TP = PROD * PP, THREADS = ka_threads(PP), NB = 2 * PP;
    __shared__ uint32_t Are[KA_DEPTH][TP][32], Aim[KA_DEPTH][TP][32], Anim[KA_DEPTH][TP][32];   // re, im, Q - im; polynomial col * 4 + row
    __shared__ uint64_t empty_bar[KA_DEPTH];                     // consumers -> producers: slot may be overwritten
    __shared__ uint32_t tws[LAB_TWS_ROWS][32];                   // forward transform constants of the 32 lanes
    __shared__ uint32_t hoist[PROD][16];                         // per producer warp: LabHoist of its current counter range
    __shared__ uint64_t wfull_bar[KA_DEPTH];                     // TMA -> consumers: witness columns of the tile have landed
    extern __shared__ __align__(128) uint32_t Sw[];              // [KA_DEPTH][COLS][KA_CONS * IC][32] witness hats of the tile
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < KA_DEPTH; s++) { mbar_init(&empty_bar[s], 32 * KA_CONS); mbar_init(&wfull_bar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 0) lab_warp_tw_to_smem(tws, lane);
    __syncthreads();
    const uint64_t rblk = (uint64_t)blockIdx.x * KA_RT;          // first row of this CTA, relative to row0
    const uint32_t ntiles = (N + COLS - 1) / COLS;
    // FULL[s] is named barrier 1 + s (producers arrive, consumers sync); EMPTY[s] is an mbarrier only the consumers
    // arrive on, so producer warps never wait for each other
    if (w < PROD) {
        // ---------------- producer ----------------
        if constexpr (PP == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        const int grow = w & 3, gcol = w >> 2;                  // polynomial p of this warp is column gcol + p * (PROD / 4) of the tile
        const bool row_ok = rblk + grow < nrows;
        // counter of coefficient 0 of A[row][n]: (row * N + n) * 64 (structs.rs:55-72); < 2^64 for every supported shape
        uint64_t ctr = ((row0 + rblk + grow) * (uint64_t)N + gcol) * 64ull;
        uint64_t tag = ~0ull;                                    // which counter range hoist[w] is valid for
        constexpr uint32_t PSTEP = 64u * (PROD / 4);             // counter distance between the warp's polynomials
        for (uint32_t t = 0; t < ntiles; t++) {
            const int s = t % KA_DEPTH;
            uint32_t c[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) c[b] = 0;
            if (row_ok && COLS * t + gcol < N) {
                const uint64_t s0 = seed.limb[0] + ctr;
                const uint64_t ntag = ((uint64_t)(s0 < ctr) << 32) | (s0 >> 32);
                if (ntag != tag) {                               // warp-uniform, once per 2^32 counters
                    LabHoist hh;
                    lab_hoist_compute(seed, ctr, 0ull, hh);
                    __syncwarp();
                    if (lane == 0) {
                        uint32_t *q = hoist[w];
                        q[0] = hh.k3; q[1] = hh.P0; q[2] = hh.P1; q[3] = hh.Q0; q[4] = hh.A5; q[5] = hh.A10; q[6] = hh.Q1; q[7] = hh.Q2;
                        q[8] = hh.A6; q[9] = hh.A2; q[10] = hh.A8; q[11] = hh.A13; q[12] = hh.A4; q[13] = hh.A9; q[14] = hh.A14;
                    }
                    __syncwarp();
                    tag = ntag;
                }
                const uint32_t lo32 = (uint32_t)s0;
                // a tile slice that straddles a 2^32 boundary of seed + counter (key word 6 differs from the hoisted one for
                // some lanes; once per 2^26 polynomials) takes the generic path, and so do coefficients whose first draw word 3 does not decide
                const bool straddle = lo32 > 0xFFFFFFFFu - (PSTEP * (PP - 1) + 63u);
                LabHoist h;
                const uint32_t *q = hoist[w];
                h.k3 = q[0]; h.P0 = q[1]; h.P1 = q[2]; h.Q0 = q[3]; h.A5 = q[4]; h.A10 = q[5]; h.Q1 = q[6]; h.Q2 = q[7];
                h.A6 = q[8]; h.A2 = q[9]; h.A8 = q[10]; h.A13 = q[11]; h.A4 = q[12]; h.A9 = q[13]; h.A14 = q[14];
                uint32_t k7[NB], w3[NB];
#pragma unroll
                for (int b = 0; b < NB; b++) k7[b] = lab_bswap32(lo32 + (uint32_t)lane + 32u * (b & 1) + PSTEP * (b >> 1));
                lab_chacha_w3<NB, RM, 1 << ((LAB_KA_VAR >> 2) & 3)>(seed, h, k7, w3);
                uint32_t slow = straddle ? (1u << NB) - 1u : 0u;
#pragma unroll
                for (int b = 0; b < NB; b++) slow |= lab_sample_w3(w3[b], c[b]) ? 0u : 1u << b;
                if (slow) {      // inlined: a call here would pin the ChaCha state to the ABI's registers
#pragma unroll
                    for (int b = 0; b < NB; b++)
                        if (slow >> b & 1u) c[b] = lab_crs_coeff_generic(seed, ctr + lane + 32u * (b & 1) + PSTEP * (b >> 1), 0ull, 0u);
                }
#pragma unroll
                for (int p = 0; p < PP; p++) {
                    if (PP > 1 && !(COLS * t + gcol + p * (PROD / 4) < N)) { c[2 * p] = 0; c[2 * p + 1] = 0; }   // column past N: zero polynomial
                    else {
                        lab_ntt32_fwd_warp_smem<(LAB_KA_VAR & 1) != 0>(c[2 * p], c[2 * p + 1], tws, lane, seed.one);
                    }
                }
            }
            ctr += 64ull * COLS;
            if (t >= KA_DEPTH) mbar_wait(&empty_bar[s], (t / KA_DEPTH - 1) & 1);  // slot free again?
#pragma unroll
            for (int p = 0; p < PP; p++) {
                const int slot = (gcol + p * (PROD / 4)) * 4 + grow;
                Are[s][slot][lane] = c[2 * p];
                Aim[s][slot][lane] = c[2 * p + 1];
                Anim[s][slot][lane] = LABQ - c[2 * p + 1];
            }
            bar_arrive_named(1 + s, THREADS);                                    // slot full
        }
    } else {
        // ---------------- consumer ----------------
        if constexpr (PP == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
        const int cw = w - PROD;
        const uint32_t i0 = i_base + (uint32_t)cw * IC;
        uint32_t accr[KA_RT][IC], acci[KA_RT][IC];
#pragma unroll
        for (int r = 0; r < KA_RT; r++)
#pragma unroll
            for (int ii = 0; ii < IC; ii++) { accr[r][ii] = 0; acci[r][ii] = 0; }
        // The transformed witness of the tile (COLS columns x 4 * IC vectors, 128 B each) is brought into a shared-memory ring
        // by TMA bulk copies that one consumer thread issues KA_DEPTH - 1 tiles ahead; the MAC loop reads it as two 16-bit
        // halves per slot (re at +0, im at +2 bytes of the packed word): no address arithmetic, no ALU-pipe unpacking, no
        // prefetch registers.  No bounds predicates either: What is padded (KA_PAD_COLS, KA_PAD_VECS) so that columns
        // n >= N and vectors i >= R are readable; the producers emit zero polynomials for n >= N and results for i >= R
        // are never stored, so whatever is read there cannot reach the output.
        constexpr uint32_t CHUNK = KA_CONS * IC * 128u;           // bytes of one column's slice
        constexpr uint32_t SLOT_WORDS = COLS * KA_CONS * IC * 32;
        const bool loader = (cw == 0 && lane == 0);
        auto load_tile = [&](uint32_t u) {                        // tile u -> ring slot u % KA_DEPTH
            const int su = u % KA_DEPTH;
            mbar_expect_tx(&wfull_bar[su], COLS * CHUNK);
#pragma unroll
            for (int nn = 0; nn < COLS; nn++)
                bulk_g2s(Sw + su * SLOT_WORDS + nn * (KA_CONS * IC * 32), What + ((size_t)(u * COLS + nn) * R + i_base) * 32, CHUNK, &wfull_bar[su]);
        };
        if (loader)
            for (uint32_t u = 0; u < ntiles && u < KA_DEPTH - 1; u++) load_tile(u);
        constexpr int FOLD_EVERY = 30 / (2 * COLS);               // tiles between folds: 2 * COLS products < 2^26 each per tile
        int pending = 0;
        const uint32_t sw_lane = (uint32_t)__cvta_generic_to_shared(Sw) + (uint32_t)(cw * IC * 32 + lane) * 4u;
        for (uint32_t t = 0; t < ntiles; t++) {
            const int s = t % KA_DEPTH;
            if (loader && t + KA_DEPTH - 1 < ntiles) {            // slot of tile t - 1 is free once every consumer has passed it
                if (t >= 1) mbar_wait(&empty_bar[(t - 1) % KA_DEPTH], ((t - 1) / KA_DEPTH) & 1);
                load_tile(t + KA_DEPTH - 1);
            }
            __syncwarp();
            const uint32_t *a_re = &Are[s][0][lane], *a_im = &Aim[s][0][lane], *a_nim = &Anim[s][0][lane];
            mbar_wait(&wfull_bar[s], (t / KA_DEPTH) & 1);         // witness slice landed
            bar_sync_named(1 + s, THREADS);                       // wait for the producers
            const uint32_t sw_slot = sw_lane + (uint32_t)s * (SLOT_WORDS * 4u);
#pragma unroll
            for (int nn = 0; nn < COLS; nn++) {
                uint32_t ar[KA_RT], am[KA_RT], nam[KA_RT];
#pragma unroll
                for (int r = 0; r < KA_RT; r++) {
                    ar[r] = a_re[(nn * 4 + r) * 32];
                    am[r] = a_im[(nn * 4 + r) * 32];
                    nam[r] = a_nim[(nn * 4 + r) * 32];
                }
#pragma unroll
                for (int ii = 0; ii < IC; ii++) {
                    const uint32_t sr = lds_u16(sw_slot + (uint32_t)(nn * KA_CONS * IC + ii) * 128u);
                    const uint32_t sm_ = lds_u16(sw_slot + (uint32_t)(nn * KA_CONS * IC + ii) * 128u + 2u);
#pragma unroll
                    for (int r = 0; r < KA_RT; r++) {         // four chained IMADs, nothing else
                        accr[r][ii] = ar[r] * sr + accr[r][ii];
                        accr[r][ii] = nam[r] * sm_ + accr[r][ii];
                        acci[r][ii] = ar[r] * sm_ + acci[r][ii];
                        acci[r][ii] = am[r] * sr + acci[r][ii];
                    }
                }
            }
            mbar_arrive(&empty_bar[s]);                          // slot may be overwritten
            if (++pending == FOLD_EVERY) {
                pending = 0;
#pragma unroll
                for (int r = 0; r < KA_RT; r++)
#pragma unroll
                    for (int ii = 0; ii < IC; ii++) { accr[r][ii] = lab_fold(accr[r][ii]); acci[r][ii] = lab_fold(acci[r][ii]); }
            }
        }
        // inverse transform and store T[i][t_row_off + row][.]  (T is [R][t_stride][64]; a row shard of a multi-GPU proof
        // writes its rows straight into the full T that the all-gather completes in place)
        const LabWarpTw tw = lab_warp_tw(lane);
#pragma unroll
        for (int ii = 0; ii < IC; ii++)
#pragma unroll
            for (int r = 0; r < KA_RT; r++) {
                uint32_t re = lab_canon(accr[r][ii]), im = lab_canon(acci[r][ii]);
                lab_ntt32_inv_warp(re, im, tw, lane);
                const uint32_t i = i0 + ii;
                if (rblk + r < nrows && i < R) {
                    uint32_t *dst = T + ((size_t)i * t_stride + t_row_off + rblk + r) * 64;
                    dst[lane] = re;
                    dst[lane + 32] = im;
                }
            }
    }
}

// ------------------------------------------------------------------------------------------------
// K_MV: generic CRS mat-vec in the transform domain (outer commitments u_1, u_2 and A*z):
//   out[x] = sum_items sum_{y in item} hat(CRS(base + x*row_stride + (y / nk)*sp + (y % nk)*sk)) (*) V[vec_off + y]
// One warp per (row x, item); partial sums go to `partial[(x * items_per_row + item)]`, k_finish_rows
// adds them, inverts and writes canonical polynomials.  (proofgen.rs:101-153, 364-378; verification.rs:274-279)
// ------------------------------------------------------------------------------------------------
struct MvItem {
    uint64_t base_lo, base_hi;   // counter of coefficient 0 of (x = 0, y = 0)
    uint64_t row_stride;         // counter step per output row x
    uint64_t sp, sk;             // counter steps of the two-level index y = p * nk + k
    uint32_t nk;
    uint32_t y0, cnt;            // this item covers y in [y0, y0 + cnt)
    uint32_t vec_off;            // V index of y = 0
    uint32_t poff;               // polynomials of this row that precede the item (index into a row of the CRS cache)
    uint32_t pad;
};

// cache_out (nullable): the generated polynomials are also written, as packed hats, to cache_out[(xr * row_polys + poff + y - y0)]
// -- the CRS cache that lets the verifier's recomputation and later proofs under the same CRS skip ChaCha20 altogether
template <bool FILL_CACHE>
__global__ void __launch_bounds__(256) k_crs_matvec(LabSeed seed, const MvItem *__restrict__ items, uint32_t items_per_row, uint64_t n_rows,
                                                    uint64_t x0, const uint32_t *__restrict__ V, uint32_t *__restrict__ partial,
                                                    uint32_t *__restrict__ cache_out, uint64_t row_polys) {
    // (hoisted ChaCha20 state and transform constants in registers, index arithmetic with divisions: both the shared-memory
    //  variant that k_gen_planes uses and an incremental index measured 3-5 % slower in this kernel)
    const int lane = threadIdx.x & 31;
    const LabWarpTw tw = lab_warp_tw(lane);
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t total = n_rows * items_per_row;
    if (wid >= total) return;
    const uint64_t xr = wid / items_per_row;
    const MvItem it = items[wid % items_per_row];
    const uint64_t x = x0 + xr;
    // base + x * row_stride as 128 bit
    uint64_t lo, hi;
    {
        unsigned __int128 b = ((unsigned __int128)it.base_hi << 64) | it.base_lo;
        b += (unsigned __int128)x * it.row_stride;
        lo = (uint64_t)b; hi = (uint64_t)(b >> 64);
    }
    uint32_t accr = 0, acci = 0;
    int pending = 0;
    LabHoist h;
    lab_hoist_invalidate(h);
    uint32_t *co = FILL_CACHE ? cache_out + (xr * row_polys + it.poff) * 32 + lane : nullptr;
    for (uint32_t y = it.y0; y < it.y0 + it.cnt; y++) {
        const uint64_t add = (uint64_t)(y / it.nk) * it.sp + (uint64_t)(y % it.nk) * it.sk;
        const uint64_t plo = lo + add;
        const uint64_t phi = hi + (plo < lo);
        uint32_t re, im;
        crs_poly_hat<LAB_RM_MATVEC>(seed, h, plo, phi, tw, lane, re, im);
        if (FILL_CACHE) { *co = lab_pack(re, im); co += 32; }
        const uint32_t v = __ldg(V + ((size_t)it.vec_off + y) * 32 + lane);
        accr += re * lab_re(v) + (LABQ - im) * lab_im(v);
        acci += re * lab_im(v) + im * lab_re(v);
        if (++pending == 16) { accr = lab_fold(accr); acci = lab_fold(acci); pending = 0; }
    }
    partial[wid * 32 + lane] = lab_pack(lab_canon(accr), lab_canon(acci));
}
// generation only: the hats of the same work items into `hats[(xr * row_polys + poff + y - y0)]`, no multiply.  The whole-proof
// graph of a small shape starts this at time zero -- the CRS does not depend on the witness -- and multiplies with
// k_cached_matvec once the digits of T and g exist, so that a default-size proof's 8.4e6 ChaCha20 blocks no longer wait for
// the inner commitment.
__global__ void __launch_bounds__(256) k_crs_gen_hats(LabSeed seed, const MvItem *__restrict__ items, uint32_t items_per_row, uint64_t n_rows, uint64_t x0,
                                                      uint32_t *__restrict__ hats, uint64_t row_polys) {
    const int lane = threadIdx.x & 31;
    const LabWarpTw tw = lab_warp_tw(lane);
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t total = n_rows * items_per_row;
    if (wid >= total) return;
    const uint64_t xr = wid / items_per_row;
    const MvItem it = items[wid % items_per_row];
    uint64_t lo, hi;
    {
        unsigned __int128 b = ((unsigned __int128)it.base_hi << 64) | it.base_lo;
        b += (unsigned __int128)(x0 + xr) * it.row_stride;
        lo = (uint64_t)b; hi = (uint64_t)(b >> 64);
    }
    LabHoist h;
    lab_hoist_invalidate(h);
    uint32_t *co = hats + (xr * row_polys + it.poff) * 32 + lane;
    for (uint32_t y = it.y0; y < it.y0 + it.cnt; y++) {
        const uint64_t add = (uint64_t)(y / it.nk) * it.sp + (uint64_t)(y % it.nk) * it.sk;
        const uint64_t plo = lo + add;
        const uint64_t phi = hi + (plo < lo);
        uint32_t re, im;
        crs_poly_hat<LAB_RM_MATVEC>(seed, h, plo, phi, tw, lane, re, im);
        *co = lab_pack(re, im);
        co += 32;
    }
}
// the same mat-vec from cached hats: HBM-bound stream of 128 B per CRS polynomial, no ChaCha20
__global__ void __launch_bounds__(256) k_cached_matvec(const uint32_t *__restrict__ cache, uint64_t row_polys, const MvItem *__restrict__ items,
                                                       uint32_t items_per_row, uint64_t n_rows, const uint32_t *__restrict__ V, uint32_t *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t total = n_rows * items_per_row;
    if (wid >= total) return;
    const uint64_t xr = wid / items_per_row;
    const MvItem it = items[wid % items_per_row];
    const uint32_t *ci = cache + (xr * row_polys + it.poff) * 32 + lane;
    const uint32_t *vi = V + ((size_t)it.vec_off + it.y0) * 32 + lane;
    uint32_t accr = 0, acci = 0;
    uint32_t y = 0;
    for (; y + 4 <= it.cnt; y += 4) {                       // four independent 128 B streams in flight per warp
        uint32_t a[4], v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { a[u] = __ldcs(ci + (size_t)(y + u) * 32); v[u] = __ldg(vi + (size_t)(y + u) * 32); }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            accr += lab_re(a[u]) * lab_re(v[u]) + (LABQ - lab_im(a[u])) * lab_im(v[u]);
            acci += lab_re(a[u]) * lab_im(v[u]) + lab_im(a[u]) * lab_re(v[u]);
        }
        if ((y & 12) == 12) { accr = lab_fold(accr); acci = lab_fold(acci); }     // every 16 terms: 16 * 2 * 2^26 < 2^32
    }
    accr = lab_fold(accr); acci = lab_fold(acci);
    for (; y < it.cnt; y++) {
        const uint32_t a = __ldcs(ci + (size_t)y * 32), v = __ldg(vi + (size_t)y * 32);
        accr += lab_re(a) * lab_re(v) + (LABQ - lab_im(a)) * lab_im(v);
        acci += lab_re(a) * lab_im(v) + lab_im(a) * lab_re(v);
    }
    partial[wid * 32 + lane] = lab_pack(lab_canon(accr), lab_canon(acci));
}

__global__ void __launch_bounds__(256) k_finish_rows(const uint32_t *__restrict__ partial, uint32_t items_per_row, uint64_t n_rows, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const LabWarpTw tw = lab_warp_tw(lane);
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (x >= n_rows) return;
    uint32_t r = 0, i = 0;
    for (uint32_t t = 0; t < items_per_row; t++) {
        uint32_t v = partial[(x * items_per_row + t) * 32 + lane];
        r += lab_re(v); i += lab_im(v);
        if ((t & 1023) == 1023) { r = lab_fold(r); i = lab_fold(i); }
    }
    r = lab_canon(r); i = lab_canon(i);
    lab_ntt32_inv_warp(r, i, tw, lane);
    out[x * 64 + lane] = r;
    out[x * 64 + 32 + lane] = i;
}
// the same with one CTA per row: its 8 warps split the items (few rows, many items per row: the outer commitments of small proofs)
__global__ void __launch_bounds__(256) k_finish_rows_cta(const uint32_t *__restrict__ partial, uint32_t items_per_row, uint64_t n_rows, uint32_t *__restrict__ out) {
    __shared__ uint32_t sre[8][32], sim[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t x = blockIdx.x;
    uint32_t r = 0, i = 0, cnt = 0;
    for (uint32_t t = w; t < items_per_row; t += 8) {
        uint32_t v = partial[(x * items_per_row + t) * 32 + lane];
        r += lab_re(v); i += lab_im(v);
        if ((++cnt & 1023) == 0) { r = lab_fold(r); i = lab_fold(i); }
    }
    sre[w][lane] = lab_canon(r);
    sim[w][lane] = lab_canon(i);
    __syncthreads();
    if (w == 0) {
        r = 0; i = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { r += sre[k][lane]; i += sim[k][lane]; }
        r = lab_canon(r); i = lab_canon(i);
        const LabWarpTw tw = lab_warp_tw(lane);
        lab_ntt32_inv_warp(r, i, tw, lane);
        out[x * 64 + lane] = r;
        out[x * 64 + 32 + lane] = i;
    }
}

// ------------------------------------------------------------------------------------------------
// JL projection and Pi^T omega: lab_jl.cuh (2-bit packed matrices, table-lookup projection)
// ------------------------------------------------------------------------------------------------
// phi''[i][n][d] = psi * phi[i][n][d] + sigma_inv(v_poly)[d]   (proofgen.rs:234-255)
__global__ void k_phi_pp(const uint32_t *__restrict__ phi, const uint32_t *__restrict__ v, uint32_t psi, size_t n_coeffs, uint32_t *__restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_coeffs) return;
    const size_t p = idx >> 6;
    const int d = (int)(idx & 63);
    uint32_t sv = v[p * 64 + ((64 - d) & 63)];
    uint32_t conj = d == 0 ? sv : (sv ? LABQ - sv : 0u);
    out[idx] = lab_canon(lab_canon(phi[idx]) * psi + conj);
}
// the same with psi read from device memory (whole-proof CUDA graph: psi changes per proof, kernel arguments do not)
__global__ void k_phi_pp_dev(const uint32_t *__restrict__ phi, const uint32_t *__restrict__ v, const uint32_t *__restrict__ psi_dev, size_t n_coeffs,
                             uint32_t *__restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_coeffs) return;
    const uint32_t psi = lab_canon(*psi_dev);
    const size_t p = idx >> 6;
    const int d = (int)(idx & 63);
    uint32_t sv = v[p * 64 + ((64 - d) & 63)];
    uint32_t conj = d == 0 ? sv : (sv ? LABQ - sv : 0u);
    out[idx] = lab_canon(lab_canon(phi[idx]) * psi + conj);
}
// int64 projection -> mod Q lift (proofgen.rs:186) is done on the host (256 values)


// ------------------------------------------------------------------------------------------------
// seeded synthetic inputs on the device (same SplitMix64 counter PRG as labrador_b200/synth.py and the
// oracle's generators): uniform Z_q values and JL matrices {-1,0,1} with P = (1/4,1/2,1/4)
// (verification.rs:553-566).  Used by bench.py and as the device-side challenge source.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t prg_u64(uint64_t base, uint64_t idx) {
    uint64_t z = base + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_synth_zq(uint64_t base, uint64_t start, size_t n, uint32_t *__restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < n; idx += stride) out[idx] = (uint32_t)__umul64hi(prg_u64(base, start + idx), (uint64_t)LABQ);
}
__global__ void k_synth_pi(uint64_t base, size_t total, int8_t *__restrict__ out) {
    size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nwords = (total + 31) / 32;
    for (; wd < nwords; wd += stride) {
        const uint64_t bits = prg_u64(base, wd);
        for (int t = 0; t < 32 && wd * 32 + t < total; t++) {
            const unsigned two = (unsigned)(bits >> (2 * t)) & 3u;
            out[wd * 32 + t] = (int8_t)(two == 0 ? -1 : (two == 3 ? 1 : 0));
        }
    }
}
// ALU-pipe ceiling: 8 independent xor+rotate chains per thread (LOP3 + SHF, the ChaCha20 quarter-round
// mix minus the adds, which can issue on the FMA pipe).  2 * 8 * iters ALU-pipe ops per thread.
__global__ void __launch_bounds__(256) k_alu_peak(uint32_t *out, int iters) {
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t t = x[j] ^ x[(j + 3) & 7];
                x[j] = __funnelshift_l(t, t, 7 + u);
            }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc ^= x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace lab
