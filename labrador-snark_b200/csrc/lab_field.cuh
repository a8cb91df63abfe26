// F_Q and F_{Q^2} arithmetic for Q = 2^13 - 1 (constants.rs:195), lazy Mersenne reduction.
// A residue is a u32 that is only congruent to its value mod Q until lab_canon() is applied.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LAB_HD __host__ __device__ __forceinline__
#define LAB_DEV __device__ __forceinline__
#else
#define LAB_HD inline
#define LAB_DEV inline
#endif

constexpr uint32_t LABQ = 8191u;
constexpr int LABD = 64;

// x == fold(x) (mod Q); fold(x) = (x & Q) + (x >> 13) <= Q + (x >> 13).  Written as x - (x >> 13) * Q so that it
// compiles to one shift (ALU pipe) + one multiply-add (FMA pipe): the ChaCha-bound kernels are limited by the ALU
// pipe and the transform kernels get an even ALU/FMA split.
LAB_HD uint32_t lab_fold(uint32_t x) { return x - (x >> 13) * LABQ; }
// x in [0, 2Q) -> [0, Q)
LAB_HD uint32_t lab_csub(uint32_t x) {
    uint32_t y = x - LABQ;
    return y < x ? y : x;
}
// any u32 -> canonical [0, Q)
LAB_HD uint32_t lab_canon(uint32_t x) { return lab_csub(lab_fold(lab_fold(x))); }
// packed complex: re | im << 16, both halves < 2^16
LAB_HD uint32_t lab_pack(uint32_t re, uint32_t im) { return re | (im << 16); }
LAB_HD uint32_t lab_re(uint32_t p) { return p & 0xFFFFu; }
LAB_HD uint32_t lab_im(uint32_t p) { return p >> 16; }

// (ar + i ai) * (br + i bi), operands < 2^14: products < 2^28, sums < 2^29. Result < 2Q.
LAB_HD void lab_cmul(uint32_t ar, uint32_t ai, uint32_t br, uint32_t bi, uint32_t &cr, uint32_t &ci) {
    uint32_t nbi = 2u * LABQ - bi;                 // -bi, positive for bi < 2^14
    uint32_t r = ar * br + ai * nbi;
    uint32_t i = ar * bi + ai * br;
    cr = lab_fold(lab_fold(r));
    ci = lab_fold(lab_fold(i));
}
