/*
 * labrador_oracle.c -- CPU restatement of the LaBRADOR prover hot path (TEST INFRASTRUCTURE).
 * See labrador_oracle.h for the parity status ("PARITY UNPINNED" for the CRS sampler bit stream).
 * Every function cites the reference lines (relative to /root/reference/) it restates.
 * Nothing here is copied from the reference: the reference is Rust over third-party crates; this
 * is plain C over dense uint32_t[64] polynomials.
 */
#define _GNU_SOURCE
#include "labrador_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

typedef unsigned __int128 u128;
#define Q LO_Q
#define D LO_D

int lo_num_threads_default(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* minimal pthread parallel-for (dynamic chunks of 1): plays the role of the reference's rayon
 * into_par_iter (proofgen.rs:101-124) for the all-cores baseline */
typedef struct {
    void (*fn)(uint64_t idx, void *ctx);
    void *ctx;
    uint64_t n;
    uint64_t next;
    pthread_mutex_t mu;
} par_job;
static void *par_worker(void *arg) {
    par_job *j = arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint64_t i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        j->fn(i, j->ctx);
    }
    return NULL;
}
static void par_for(uint64_t n, int nthreads, void (*fn)(uint64_t, void *), void *ctx) {
    if (nthreads <= 1 || n <= 1) { for (uint64_t i = 0; i < n; i++) fn(i, ctx); return; }
    if ((uint64_t)nthreads > n) nthreads = (int)n;
    par_job j = { fn, ctx, n, 0, PTHREAD_MUTEX_INITIALIZER };
    pthread_t *th = malloc((size_t)nthreads * sizeof *th);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, par_worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------------
 * constants.rs:234-264  RuntimeConstants::new -- same f64 operation order; Rust's f64 methods
 * lower to the platform libm (glibc here and on the GPU box), `as i128` saturates.
 * ------------------------------------------------------------------------------------------ */
static int64_t sat_i64(double x, int *bad) {
    if (!(x == x)) { *bad = 1; return 0; }                 /* NaN as i128 == 0 in Rust */
    if (x >= 9.2e18) { *bad = 1; return INT64_MAX; }       /* reference would hold i128::MAX */
    if (x <= -9.2e18) { *bad = 1; return INT64_MIN; }
    return (int64_t)x;                                     /* truncation toward zero, like `as` */
}

int lo_runtime_constants(uint64_t N, uint64_t R, lo_constants *o) {
    int bad = 0;
    double q = (double)Q;
    memset(o, 0, sizeof *o);
    o->N = N; o->R = R;
    o->KAPPA = N * D; o->KAPPA_1 = N * D; o->KAPPA_2 = N * D;              /* :237-239 */
    o->BETA_BOUND = sat_i64(floor(sqrt(30.0 / 128.0) * q / 125.0), &bad); /* :241 */
    o->STD = (double)o->BETA_BOUND / sqrt((double)(R * N * D));           /* :242 */
    o->B = sat_i64(round(sqrt(sqrt(12. * (double)R * LO_TAU) * o->STD)), &bad); /* :243 */
    o->T_1 = sat_i64(round(log2(q) / log2((double)o->B)), &bad);          /* :244 */
    o->B_1 = sat_i64(pow(q, 1.0 / (double)o->T_1), &bad);                 /* :245 (truncates) */
    double nd = (double)(N * D);
    o->T_2 = sat_i64(round(log2(sqrt(24. * nd) * (o->STD * o->STD)) / log2((double)o->B)), &bad); /* :246 */
    o->B_2 = sat_i64(round(pow(sqrt((double)(24 * (N * D))) * (o->STD * o->STD), 1.0 / (double)o->T_2)), &bad); /* :247 */
    double b1 = (double)o->B_1, b2 = (double)o->B_2, t1 = (double)o->T_1, t2 = (double)o->T_2;
    double r = (double)R;
    o->GAMMA = (double)(o->BETA_BOUND * o->BETA_BOUND) * LO_TAU;           /* :250 */
    o->GAMMA_1 = ((b1 * b1 * t1) / 12.0) * r * (double)o->KAPPA * (double)D
               + ((b2 * b2 * t2) / 12.0) * ((r * r + r) / 2.0) * (double)D; /* :251 */
    o->GAMMA_2 = ((b1 * b1 * t1) / 12.0) * ((r * r + r) / 2.0) * (double)D; /* :252 */
    double bb = (double)o->B;
    o->BETA_PRIME = (2.0 / (bb * bb)) * o->GAMMA + o->GAMMA_1 + o->GAMMA_2; /* :254 */
    if (bad || o->B < 2 || o->T_1 <= 0 || o->B_1 < 2 || o->T_2 <= 0 || o->B_2 < 2 ||
        o->T_1 > 64 || o->T_2 > 64)
        o->degenerate = 1;
    return o->degenerate ? LO_ERR_PARAMS : LO_OK;
}

/* ------------------------------------------------------------------------------------------
 * Z_q / R_q
 * ------------------------------------------------------------------------------------------ */
uint32_t lo_mod_positive(int64_t v) {            /* util.rs:16-23 */
    int64_t r = v % (int64_t)Q;
    return (uint32_t)(r < 0 ? r + (int64_t)Q : r);
}

/* algebraic.rs:379-404 with NTT off: schoolbook product (degree <= 126), then Rq::reduction
 * folds the term of degree deg >= D onto deg % D with sign (-1)^(deg / D) (:364-374). */
void lo_rq_mul(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    int64_t acc[2 * D];
    memset(acc, 0, sizeof acc);
    for (int i = 0; i < D; i++) {
        int64_t ai = a[i];
        if (!ai) continue;
        for (int j = 0; j < D; j++) acc[i + j] += ai * (int64_t)b[j];
    }
    for (int d = 0; d < D; d++) c[d] = lo_mod_positive(acc[d] - acc[d + D]);
}
void lo_rq_add(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    for (int d = 0; d < D; d++) { uint32_t s = a[d] + b[d]; c[d] = s >= Q ? s - Q : s; }
}
void lo_rq_sub(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    for (int d = 0; d < D; d++) { uint32_t s = a[d] + Q - b[d]; c[d] = s >= Q ? s - Q : s; }
}
void lo_rq_scale(const uint32_t *a, uint32_t s, uint32_t *c) {   /* util.rs:176-180 */
    for (int d = 0; d < D; d++) c[d] = (uint32_t)(((uint64_t)a[d] * s) % Q);
}
/* util.rs:496-509 */
void lo_inner_product(const uint32_t *v1, const uint32_t *v2, size_t n, uint32_t *out) {
    uint32_t acc[D], prod[D];
    memset(acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) {
        lo_rq_mul(v1 + i * D, v2 + i * D, prod);
        lo_rq_add(acc, prod, acc);
    }
    memcpy(out, acc, sizeof acc);
}
/* util.rs:118-137: c0 stays, coefficient of X^d goes to X^(D-d) negated */
void lo_sigma_inv(const uint32_t *a, uint32_t *out) {
    uint32_t tmp[D];
    memset(tmp, 0, sizeof tmp);
    for (int d = 0; d < D; d++) {
        if (d == 0) tmp[0] = a[0];
        else tmp[D - d] = a[d] ? Q - a[d] : 0;   /* Neg for Zq: algebraic.rs:56-63 */
    }
    memcpy(out, tmp, sizeof tmp);
}
uint64_t lo_norm_sq(const uint32_t *coeffs, size_t n) {           /* util.rs:195-202 */
    uint64_t s = 0;
    for (size_t i = 0; i < n; i++) s += (uint64_t)coeffs[i] * coeffs[i];
    return s;
}

/* util.rs:389-442 executed literally with the Zq operator semantics of algebraic.rs:
 *   `poly_coeff != 0`            PartialEq<i128>  (:287-291)
 *   `poly_coeff % base`          Rem<i128> => Zq::new(value % base)  (:75-81)
 *   centered_rep (util.rs:377-387): `val > b/2` is PartialOrd<i128> against mod_positive(b/2,Q) (:293-297)
 *   `(poly_coeff - remainder)`   Sub => mod_positive  (:111-118)
 *   `/ base`                     Div<i128> => integer division of the representative (:136-142)
 * Returns LO_ERR_PARAMS if the loop cannot terminate (base < 2). */
int lo_decompose_literal(const uint32_t *p, int64_t base, int64_t exp, uint32_t *out) {
    if (base < 2 || exp <= 0) return LO_ERR_PARAMS;
    memset(out, 0, (size_t)exp * D * sizeof(uint32_t));
    for (int deg = 0; deg < D; deg++) {
        int64_t v = p[deg];
        int64_t k = 0;
        while (v != 0) {
            int64_t r0 = lo_mod_positive(v % base);
            int64_t rem;
            if (r0 > (int64_t)lo_mod_positive(base / 2)) rem = lo_mod_positive((base - r0) % base);
            else rem = r0;
            if (k < exp) out[k * D + deg] = (uint32_t)rem;   /* digits beyond exp are dropped (:425-440) */
            k++;
            v = (int64_t)lo_mod_positive(v - rem) / base;
        }
    }
    return LO_OK;
}
/* closed form: digit_k(c) = f(floor(c / b^k) mod b), f(x) = x if x <= floor(b/2) else b - x */
void lo_decompose(const uint32_t *p, int64_t base, int64_t exp, uint32_t *out) {
    for (int deg = 0; deg < D; deg++) {
        int64_t v = p[deg];
        for (int64_t k = 0; k < exp; k++) {
            int64_t x = v % base;
            out[k * D + deg] = (uint32_t)(x <= base / 2 ? x : base - x);
            v /= base;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * F_{Q^2} transform (baseline speed path + independent check of the GPU NTT design).
 * R_q = F_Q[X]/(X^64+1) ~ F_{Q^2}[X]/(X^32 - i), f -> g_d = f_d + i f_{d+32}; X^32 - i splits
 * over F_{Q^2} = F_Q[i] because 128 | Q^2-1.  zeta = 2620 + 936 i, zeta^32 = i.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t re, im; } cq;
static inline cq cq_mul(cq x, cq y) {
    cq r;
    r.re = (uint32_t)(((uint64_t)x.re * y.re + (uint64_t)(Q - x.im) * y.im) % Q);
    r.im = (uint32_t)(((uint64_t)x.re * y.im + (uint64_t)x.im * y.re) % Q);
    return r;
}
static inline cq cq_add(cq x, cq y) { cq r = { (x.re + y.re) % Q, (x.im + y.im) % Q }; return r; }
static inline cq cq_sub(cq x, cq y) { cq r = { (x.re + Q - y.re) % Q, (x.im + Q - y.im) % Q }; return r; }
static cq cq_pow(cq x, uint64_t e) {
    cq r = { 1, 0 };
    while (e) { if (e & 1) r = cq_mul(r, x); x = cq_mul(x, x); e >>= 1; }
    return r;
}
static const cq ZETA = { 2620, 936 };
static cq tw_fwd[32], tw_inv[32];
static int tw_exp[32], slot_exp[32];
static int ntt_ready = 0;
static void ntt_init(void) {
    if (ntt_ready) return;
    /* node k (k=1 root) has modulus X^(2len) - zeta^e; butterfly twiddle zeta^(e/2);
       children 2k: e/2, 2k+1: e/2 + 64 */
    int node_e[64];
    node_e[1] = 32;
    for (int k = 1; k < 32; k++) {
        tw_exp[k] = node_e[k] / 2;
        node_e[2 * k] = node_e[k] / 2;
        node_e[2 * k + 1] = node_e[k] / 2 + 64;
        tw_fwd[k] = cq_pow(ZETA, (uint64_t)tw_exp[k]);
        tw_inv[k] = cq_pow(ZETA, (uint64_t)(128 - tw_exp[k]));
    }
    for (int j = 0; j < 32; j++) slot_exp[j] = node_e[32 + j];
    ntt_ready = 1;
}
int lo_ntt_slot_exponent(int j) { ntt_init(); return slot_exp[j]; }
void lo_ntt_zeta(uint32_t *re, uint32_t *im) { *re = ZETA.re; *im = ZETA.im; }

void lo_ntt_fwd(const uint32_t *poly, uint32_t *re_im) {
    ntt_init();
    cq g[32];
    for (int d = 0; d < 32; d++) { g[d].re = poly[d]; g[d].im = poly[d + 32]; }
    int k = 1;
    for (int len = 16; len >= 1; len >>= 1)
        for (int start = 0; start < 32; start += 2 * len) {
            cq r = tw_fwd[k++];
            for (int j = start; j < start + len; j++) {
                cq t = cq_mul(r, g[j + len]);
                g[j + len] = cq_sub(g[j], t);
                g[j] = cq_add(g[j], t);
            }
        }
    for (int j = 0; j < 32; j++) { re_im[2 * j] = g[j].re; re_im[2 * j + 1] = g[j].im; }
}
void lo_ntt_inv(const uint32_t *re_im, uint32_t *poly) {
    ntt_init();
    cq g[32];
    for (int j = 0; j < 32; j++) { g[j].re = re_im[2 * j]; g[j].im = re_im[2 * j + 1]; }
    for (int len = 1; len <= 16; len <<= 1) {
        int k = 16 / len;                       /* first node of this level */
        for (int start = 0; start < 32; start += 2 * len) {
            cq r = tw_inv[k++];
            for (int j = start; j < start + len; j++) {
                cq u = g[j], v = g[j + len];
                g[j] = cq_add(u, v);
                g[j + len] = cq_mul(cq_sub(u, v), r);
            }
        }
    }
    for (int d = 0; d < 32; d++) {              /* 32^-1 = 2^-5 = 2^8 mod (2^13-1) */
        poly[d] = (uint32_t)(((uint64_t)g[d].re * 256) % Q);
        poly[d + 32] = (uint32_t)(((uint64_t)g[d].im * 256) % Q);
    }
}
void lo_rq_mul_ntt(const uint32_t *a, const uint32_t *b, uint32_t *c) {
    uint32_t fa[64], fb[64];
    lo_ntt_fwd(a, fa); lo_ntt_fwd(b, fb);
    for (int j = 0; j < 32; j++) {
        cq x = { fa[2 * j], fa[2 * j + 1] }, y = { fb[2 * j], fb[2 * j + 1] };
        cq z = cq_mul(x, y);
        fa[2 * j] = z.re; fa[2 * j + 1] = z.im;
    }
    lo_ntt_inv(fa, c);
}

/* ------------------------------------------------------------------------------------------
 * ChaCha20 (rand_chacha 0.3.1: 20 rounds, 64-bit block counter in words 12-13, 64-bit stream
 * in words 14-15, key = seed bytes as little-endian u32 words) [3P, restated from the
 * published ChaCha specification; KAT-pinned in tests/test_oracle_kat.py]
 * ------------------------------------------------------------------------------------------ */
#define ROTL(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
#define QR(a, b, c, d) \
    a += b; d ^= a; d = ROTL(d, 16); c += d; b ^= c; b = ROTL(b, 12); \
    a += b; d ^= a; d = ROTL(d, 8);  c += d; b ^= c; b = ROTL(b, 7);
void lo_chacha20_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]) {
    uint32_t s[16], x[16];
    s[0] = 0x61707865; s[1] = 0x3320646e; s[2] = 0x79622d32; s[3] = 0x6b206574;
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32);
    s[14] = (uint32_t)stream;  s[15] = (uint32_t)(stream >> 32);
    memcpy(x, s, sizeof x);
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

/* 256-bit big-endian base seed + 128-bit offset -> 32 key bytes (structs.rs:47-53,59-68,155-165:
 * BigUint add and the byte-wise big-endian increment agree while no carry leaves 256 bits) */
static void seed_add(const uint8_t seed[32], u128 off, uint8_t out[32]) {
    unsigned carry = 0;
    for (int i = 31; i >= 0; i--) {
        unsigned add = (unsigned)(off & 0xff);
        off >>= 8;
        unsigned s = seed[i] + add + carry;
        out[i] = (uint8_t)s;
        carry = s >> 8;
    }
}

/* structs.rs:167-171: ChaCha20Rng::from_seed(offset_seed) then gen_range(0..Q) on i128.
 * rand 0.8.5 UniformInt<i128>::sample_single (range = Q as u128, zone = (Q << lz(Q)) - 1 with
 * lz = 115): loop { v = u128 from two next_u64 (low first, each two consecutive keystream u32s,
 * low first); (hi,lo) = v * Q as 256 bit; if lo <= zone return hi }.  [3P, PARITY UNPINNED] */
uint32_t lo_crs_coeff(const uint8_t seed[32], u128 ctr) {
    uint8_t kb[32];
    uint32_t key[8], ks[16];
    seed_add(seed, ctr, kb);
    for (int i = 0; i < 8; i++)
        key[i] = (uint32_t)kb[4 * i] | ((uint32_t)kb[4 * i + 1] << 8) | ((uint32_t)kb[4 * i + 2] << 16) | ((uint32_t)kb[4 * i + 3] << 24);
    const u128 zone = (((u128)Q) << 115) - 1;
    uint64_t word = 0;       /* index of the next unread keystream u32 */
    uint64_t cur_block = ~0ull;
    for (;;) {
        uint32_t w[4];
        for (int t = 0; t < 4; t++, word++) {
            if (word / 16 != cur_block) { cur_block = word / 16; lo_chacha20_block(key, cur_block, 0, ks); }
            w[t] = ks[word % 16];
        }
        u128 v = ((u128)w[3] << 96) | ((u128)w[2] << 64) | ((u128)w[1] << 32) | (u128)w[0];
        /* v * Q as (hi, lo) */
        u128 vlo = (uint64_t)v, vhi = v >> 64;
        u128 p0 = vlo * Q;               /* < 2^77 */
        u128 p1 = vhi * Q + (p0 >> 64);  /* < 2^78 */
        u128 lo = (p1 << 64) | (uint64_t)p0;
        u128 hi = p1 >> 64;
        if (lo <= zone) return (uint32_t)hi;
    }
}
void lo_crs_poly(const uint8_t seed[32], u128 start, uint32_t *out) {   /* structs.rs:35-45 */
    for (int d = 0; d < D; d++) out[d] = lo_crs_coeff(seed, start + (u128)d);
}

/* structs.rs:55-72 */
u128 lo_off_A(const lo_constants *c, uint64_t row) { return (u128)row * (u128)(c->N * D); }
/* structs.rs:74-88 -- size_B_mat = KAPPA_1*KAPPA (no *D): consecutive B_ik overlap, literal */
u128 lo_off_B(const lo_constants *c, uint64_t i, uint64_t k, uint64_t row) {
    u128 offA = (u128)(c->KAPPA * c->N * D);
    u128 sizeB = (u128)(c->KAPPA_1 * c->KAPPA);
    return offA + (u128)(i * (uint64_t)c->T_1 + k) * sizeB + (u128)row * (u128)(c->KAPPA * D);
}
static uint64_t sum_pairs(const lo_constants *c, uint64_t i) { return i > 0 ? i * c->R - i * (i - 1) / 2 : 0; }
/* structs.rs:90-114 -- stride T_1 even though k ranges over T_2, literal */
u128 lo_off_C(const lo_constants *c, uint64_t i, uint64_t j, uint64_t k) {
    u128 offA = (u128)(c->KAPPA * c->N * D);
    u128 sizeB = (u128)(c->KAPPA_1 * c->KAPPA);
    u128 numB = (u128)(c->R * (uint64_t)c->T_1);
    u128 offC = (u128)(k + (uint64_t)c->T_1 * (sum_pairs(c, i) + (j - i)));
    return offA + numB * sizeB * D + offC * (u128)(c->KAPPA_2 * D);
}
/* structs.rs:116-144 -- num_C_matrices = R(R+1)/2 without the T factor: D overlaps C, literal */
u128 lo_off_D(const lo_constants *c, uint64_t i, uint64_t j, uint64_t k) {
    u128 offA = (u128)(c->KAPPA * c->N * D);
    u128 sizeB = (u128)(c->KAPPA_1 * c->KAPPA);
    u128 numB = (u128)(c->R * (uint64_t)c->T_1);
    u128 numC = (u128)(c->R * (c->R + 1) / 2);
    u128 offD = (u128)(k + (uint64_t)c->T_1 * (sum_pairs(c, i) + (j - i)));
    return offA + numB * sizeB * D + numC * (u128)(c->KAPPA_2 * D) + offD * (u128)(c->KAPPA_2 * D);
}
void lo_off_split(int which, const lo_constants *c, uint64_t i, uint64_t j, uint64_t k, uint64_t row, uint64_t *lo, uint64_t *hi) {
    u128 o = which == 0 ? lo_off_A(c, row) : which == 1 ? lo_off_B(c, i, k, row) : which == 2 ? lo_off_C(c, i, j, k) : lo_off_D(c, i, j, k);
    *lo = (uint64_t)o; *hi = (uint64_t)(o >> 64);
}
static void fetch_n(const uint8_t seed[32], u128 start, uint64_t n, uint32_t *out) {  /* structs.rs:147-153 */
    for (uint64_t p = 0; p < n; p++) lo_crs_poly(seed, start + (u128)p * D, out + p * D);
}
void lo_fetch_A_row(const lo_constants *c, const uint8_t seed[32], uint64_t row, uint32_t *out) { fetch_n(seed, lo_off_A(c, row), c->N, out); }
void lo_fetch_B_ik_row(const lo_constants *c, const uint8_t seed[32], uint64_t i, uint64_t k, uint64_t row, uint32_t *out) { fetch_n(seed, lo_off_B(c, i, k, row), c->KAPPA, out); }
void lo_fetch_C_ijk(const lo_constants *c, const uint8_t seed[32], uint64_t i, uint64_t j, uint64_t k, uint32_t *out) { fetch_n(seed, lo_off_C(c, i, j, k), c->KAPPA_2, out); }
void lo_fetch_D_ijk(const lo_constants *c, const uint8_t seed[32], uint64_t i, uint64_t j, uint64_t k, uint32_t *out) { fetch_n(seed, lo_off_D(c, i, j, k), c->KAPPA_2, out); }

/* ------------------------------------------------------------------------------------------
 * synthetic-input PRG (not part of the reference: the reference is unseeded)
 * ------------------------------------------------------------------------------------------ */
uint64_t lo_prg_u64(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint64_t z = seed + stream * 0xD1342543DE82EF95ull + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint32_t lo_prg_zq(uint64_t seed, uint64_t stream, uint64_t idx) {
    return (uint32_t)(((u128)lo_prg_u64(seed, stream, idx) * Q) >> 64);
}

/* ------------------------------------------------------------------------------------------
 * stages
 * ------------------------------------------------------------------------------------------ */
typedef void (*mulfn)(const uint32_t *, const uint32_t *, uint32_t *);
static void ip(mulfn mul, const uint32_t *v1, const uint32_t *v2, size_t n, uint32_t *out) {
    uint32_t acc[D], prod[D];
    memset(acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) { mul(v1 + i * D, v2 + i * D, prod); lo_rq_add(acc, prod, acc); }
    memcpy(out, acc, sizeof acc);
}

/* proofgen.rs:41-49: for each row kappa: fetch_A_row, then for each i: t_i[kappa] = <A_row, s_i> */
typedef struct { const lo_constants *c; const uint8_t *seed; const uint32_t *S; uint64_t row0, nrows; mulfn mul; uint32_t *T; } ci_ctx;
static void ci_row(uint64_t rr, void *vp) {
    ci_ctx *x = vp;
    const lo_constants *c = x->c;
    uint32_t *row = malloc(c->N * D * sizeof(uint32_t));
    lo_fetch_A_row(c, x->seed, x->row0 + rr, row);
    for (uint64_t i = 0; i < c->R; i++)
        ip(x->mul, row, x->S + i * c->N * D, c->N, x->T + (i * x->nrows + rr) * D);
    free(row);
}
void lo_commit_inner_rows(const lo_constants *c, const uint8_t seed[32], const uint32_t *S,
                          uint64_t row0, uint64_t nrows, int use_ntt, int nthreads, uint32_t *T) {
    ntt_init();
    ci_ctx x = { c, seed, S, row0, nrows, use_ntt ? lo_rq_mul_ntt : lo_rq_mul, T };
    par_for(nrows, nthreads, ci_row, &x);
}
/* proofgen.rs:59-70: all R^2 pairs */
void lo_gram(const lo_constants *c, const uint32_t *S, uint32_t *G) {
    for (uint64_t i = 0; i < c->R; i++)
        for (uint64_t j = 0; j < c->R; j++)
            lo_inner_product(S + i * c->N * D, S + j * c->N * D, c->N, G + (i * c->R + j) * D);
}
/* proofgen.rs:429-457 + util.rs:446-467,511-526: p = sum_i Pi_i * coeffs(s_i), exact integers */
void lo_jl_project(const lo_constants *c, const uint32_t *S, const int8_t *Pi, int64_t *p) {
    uint64_t nd = c->N * D;
    for (int j = 0; j < LO_JL_ROWS; j++) p[j] = 0;
    for (uint64_t i = 0; i < c->R; i++)
        for (int j = 0; j < LO_JL_ROWS; j++) {
            const int8_t *row = Pi + (i * LO_JL_ROWS + (uint64_t)j) * nd;
            const uint32_t *s = S + i * nd;
            int64_t acc = 0;
            for (uint64_t x = 0; x < nd; x++) acc += (int64_t)row[x] * (int64_t)s[x];
            p[j] += acc;
        }
}
/* verification.rs:568-579 + util.rs:216-219, literal f64 */
int lo_valid_projection(const lo_constants *c, const int64_t *p) {
    __int128 ss = 0;
    for (int j = 0; j < LO_JL_ROWS; j++) ss += (__int128)p[j] * p[j];
    double norm = sqrt((double)ss);
    return norm <= sqrt(128.) * (double)c->BETA_BOUND;
}
/* proofgen.rs:380-399 */
void lo_amortize_z(const lo_constants *c, const uint32_t *S, const uint32_t *ch, uint32_t *z) {
    uint32_t prod[D];
    memset(z, 0, c->N * D * sizeof(uint32_t));
    for (uint64_t i = 0; i < c->R; i++)
        for (uint64_t n = 0; n < c->N; n++) {
            lo_rq_mul(ch + i * D, S + (i * c->N + n) * D, prod);
            lo_rq_add(z + n * D, prod, z + n * D);
        }
}

typedef struct { const lo_constants *c; const uint8_t *seed; const uint32_t *tdec; mulfn mul; uint32_t *lhs; } ol_ctx;
static void ol_row(uint64_t x, void *vp) {
    ol_ctx *o = vp;
    const lo_constants *c = o->c;
    uint64_t R = c->R, K = c->KAPPA, T1 = (uint64_t)c->T_1;
    uint32_t *row = malloc(K * D * sizeof(uint32_t));
    uint32_t acc[D], prod[D];
    memset(acc, 0, sizeof acc);
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t k = 0; k < T1; k++) {
            lo_fetch_B_ik_row(c, o->seed, i, k, x, row);
            ip(o->mul, row, o->tdec + (i * T1 + k) * K * D, K, prod);
            lo_rq_add(acc, prod, acc);
        }
    memcpy(o->lhs + x * D, acc, sizeof acc);
    free(row);
}
/* Sum_{i,k} B_ik * dig_k(t_i)  (proofgen.rs:101-132 / verification.rs:372-396) */
static void outer_lhs(const lo_constants *c, const uint8_t seed[32], const uint32_t *T, mulfn mul, int nthreads, uint32_t *lhs) {
    uint64_t R = c->R, K = c->KAPPA, K1 = c->KAPPA_1, T1 = (uint64_t)c->T_1;
    /* decompose_polynomial_vec(t_i, B_1, T_1): tdec[i][k][y] */
    uint32_t *tdec = malloc(R * T1 * K * D * sizeof(uint32_t));
    uint32_t *dig = malloc(T1 * D * sizeof(uint32_t));
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t y = 0; y < K; y++) {
            lo_decompose(T + (i * K + y) * D, c->B_1, c->T_1, dig);
            for (uint64_t k = 0; k < T1; k++) memcpy(tdec + ((i * T1 + k) * K + y) * D, dig + k * D, D * sizeof(uint32_t));
        }
    free(dig);
    memset(lhs, 0, K1 * D * sizeof(uint32_t));
    ol_ctx x = { c, seed, tdec, mul, lhs };
    par_for(K1, nthreads, ol_row, &x);
    free(tdec);
}
/* Sum_{i<=j,k<nk} dig_k^{base}(M_ij) * X_ijk  with X = C (which=2) or D (which=3)
 * (proofgen.rs:137-148, :364-378 / verification.rs:398-408, :421-431) */
static void outer_rank1(const lo_constants *c, const uint8_t seed[32], const uint32_t *M, int which,
                        int64_t base, int64_t nk, mulfn mul, uint32_t *acc /* [KAPPA_2][64], accumulated into */) {
    uint64_t R = c->R, K2 = c->KAPPA_2;
    uint32_t *vec = malloc(K2 * D * sizeof(uint32_t));
    uint32_t *dig = malloc((size_t)nk * D * sizeof(uint32_t));
    uint32_t prod[D];
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t j = i; j < R; j++) {
            lo_decompose(M + (i * R + j) * D, base, nk, dig);
            for (int64_t k = 0; k < nk; k++) {
                if (which == 2) lo_fetch_C_ijk(c, seed, i, j, (uint64_t)k, vec);
                else lo_fetch_D_ijk(c, seed, i, j, (uint64_t)k, vec);
                for (uint64_t x = 0; x < K2; x++) {
                    mul(dig + k * D, vec + x * D, prod);
                    lo_rq_add(acc + x * D, prod, acc + x * D);
                }
            }
        }
    free(vec); free(dig);
}

/* phi''_i = psi*phi_i + sum_j omega_j * sigma_inv(pi_i^(j))   (proofgen.rs:234-255 / verification.rs:60-89) */
static void phi_prime_prime(const lo_constants *c, const lo_state *st, const int8_t *pi, uint32_t psi,
                            const uint32_t *omega, uint32_t *out /*[R][N][64]*/) {
    uint64_t R = c->R, N = c->N, nd = N * D;
    uint32_t lifted[D], conj[D], sc[D];
    for (uint64_t i = 0; i < R; i++) {
        for (uint64_t n = 0; n < N; n++)   /* multiply_poly_vec_ints with L=1 == scale by psi (util.rs:139-155) */
            lo_rq_scale(st->phi + (i * N + n) * D, psi, out + (i * N + n) * D);
        for (int j = 0; j < LO_JL_ROWS; j++) {
            const int8_t *row = pi + (i * LO_JL_ROWS + (uint64_t)j) * nd;
            for (uint64_t n = 0; n < N; n++) {
                for (int d = 0; d < D; d++) lifted[d] = lo_mod_positive(row[n * D + d]);  /* Zq::lift, proofgen.rs:445-453 */
                lo_sigma_inv(lifted, conj);
                lo_rq_scale(conj, omega[j], sc);
                lo_rq_add(out + (i * N + n) * D, sc, out + (i * N + n) * D);
            }
        }
    }
}

int lo_prove(const lo_constants *c, const uint8_t seed[32], const uint32_t *S, const lo_state *st,
             const lo_challenges *ch, int use_ntt, int nthreads, lo_transcript *out) {
    if (c->degenerate) return LO_ERR_PARAMS;
    mulfn mul = use_ntt ? lo_rq_mul_ntt : lo_rq_mul;
    ntt_init();
    uint64_t R = c->R, N = c->N, K = c->KAPPA, K1 = c->KAPPA_1, K2 = c->KAPPA_2, nd = N * D;
    uint32_t prod[D], tmp[D];

    /* S1 inner commitments (:35-49) */
    lo_commit_inner_rows(c, seed, S, 0, K, use_ntt, nthreads, out->t);
    /* S2 g (:59-70) */
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t j = 0; j < R; j++)
            ip(mul, S + i * nd, S + j * nd, N, out->g + (i * R + j) * D);
    /* S3 u_1 (:101-153) */
    outer_lhs(c, seed, out->t, mul, nthreads, out->u_1);
    {
        uint32_t *rhs = calloc(K2 * D, sizeof(uint32_t));
        outer_rank1(c, seed, out->g, 2, c->B_2, c->T_2, mul, rhs);
        for (uint64_t x = 0; x < K1; x++) lo_rq_add(out->u_1 + x * D, rhs + x * D, out->u_1 + x * D);
        free(rhs);
    }
    /* S4 JL with retries (:161-186): initial attempt + up to 5 retries, 6th rejection panics */
    int att = 0, rejections = 0;
    const int8_t *pi = ch->pi;
    for (;;) {
        if (att >= ch->n_attempts) return LO_ERR_JL_REJECTED;
        pi = ch->pi + (size_t)att * R * LO_JL_ROWS * nd;
        lo_jl_project(c, S, pi, out->projection_int);
        if (lo_valid_projection(c, out->projection_int)) break;
        rejections++;
        if (rejections > 5) return LO_ERR_JL_REJECTED;
        att++;
    }
    out->jl_attempt = att;
    for (int j = 0; j < LO_JL_ROWS; j++) out->projection[j] = lo_mod_positive(out->projection_int[j] % (int64_t)Q);
    /* S5 aggregation, upper_bound = 1 (:189-289) */
    uint32_t *phipp = malloc(R * nd * sizeof(uint32_t));
    phi_prime_prime(c, st, pi, ch->psi, ch->omega, phipp);
    uint32_t bpp[D];
    memset(bpp, 0, sizeof bpp);
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t j = 0; j < R; j++) {
            uint32_t app[D], gij[D];
            lo_rq_scale(st->a + (i * R + j) * D, ch->psi, app);          /* a''_ij (:227-231) */
            ip(mul, S + i * nd, S + j * nd, N, gij);                     /* recomputed (:261-264) */
            mul(app, gij, prod);
            lo_rq_add(bpp, prod, bpp);
        }
    for (uint64_t i = 0; i < R; i++) {
        ip(mul, phipp + i * nd, S + i * nd, N, tmp);
        lo_rq_add(bpp, tmp, bpp);
    }
    memcpy(out->b_prime_prime, bpp, sizeof bpp);
    {   /* verify_b_prime_prime (verification.rs:532-551) */
        uint64_t acc = 0;
        for (int j = 0; j < LO_JL_ROWS; j++) acc = (acc + (uint64_t)ch->omega[j] * out->projection[j]) % Q;
        uint64_t check = (acc + (uint64_t)ch->psi * st->b[0]) % Q;       /* b'_0 = b.eval(0) (structs.rs:373) */
        if (bpp[0] != check) { free(phipp); return LO_ERR_BPP_CHECK; }
    }
    /* S6 phi_final = alpha*phi + beta*phi'' (:295-314) */
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t n = 0; n < N; n++) {
            mul(ch->alpha, st->phi + (i * N + n) * D, prod);
            mul(ch->beta, phipp + (i * N + n) * D, tmp);
            lo_rq_add(prod, tmp, out->phi_final + (i * N + n) * D);
        }
    free(phipp);
    /* S7 h_ij = (<phi_i,s_j> + <phi_j,s_i>) * 2^-1, 2^-1 = 2^(Q-2) = 4096 (:320-358); the suspended
       modulus is value-neutral */
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t j = 0; j < R; j++) {
            uint32_t x1[D], x2[D];
            ip(mul, out->phi_final + i * nd, S + j * nd, N, x1);
            ip(mul, out->phi_final + j * nd, S + i * nd, N, x2);
            lo_rq_add(x1, x2, x1);
            lo_rq_scale(x1, 4096, out->h + (i * R + j) * D);
        }
    /* S8 u_2 (:364-378) */
    memset(out->u_2, 0, K2 * D * sizeof(uint32_t));
    outer_rank1(c, seed, out->h, 3, c->B_1, c->T_1, mul, out->u_2);
    /* S9 z (:380-399) */
    memset(out->z, 0, nd * sizeof(uint32_t));
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t n = 0; n < N; n++) {
            mul(ch->c + i * D, S + (i * N + n) * D, prod);
            lo_rq_add(out->z + n * D, prod, out->z + n * D);
        }
    return LO_OK;
}

/* verification.rs:25-438 */
int lo_verify(const lo_constants *c, const uint8_t seed[32], const lo_state *st, const lo_challenges *ch,
              const lo_transcript *tr, int use_ntt, int nthreads, int *failed_check, uint64_t *norm_sum) {
    mulfn mul = use_ntt ? lo_rq_mul_ntt : lo_rq_mul;
    ntt_init();
    uint64_t R = c->R, N = c->N, K = c->KAPPA, K1 = c->KAPPA_1, K2 = c->KAPPA_2, nd = N * D;
    uint32_t prod[D], tmp[D];
    int fc = 0, ok = 1;
    const int8_t *pi = ch->pi + (size_t)tr->jl_attempt * R * LO_JL_ROWS * nd;
    uint32_t *phipp = malloc(R * nd * sizeof(uint32_t));
    uint32_t *phi = malloc(R * nd * sizeof(uint32_t));
    uint32_t *acon = malloc(R * R * D * sizeof(uint32_t));
    uint32_t *cand = NULL, *rhs = NULL;
    /* lines 3-4 */
    phi_prime_prime(c, st, pi, ch->psi, ch->omega, phipp);
    /* line 5: a_ij = alpha*a_ij + beta*(psi*a_ij) */
    for (uint64_t e = 0; e < R * R; e++) {
        uint32_t app[D];
        mul(ch->alpha, st->a + e * D, prod);
        lo_rq_scale(st->a + e * D, ch->psi, app);
        mul(ch->beta, app, tmp);
        lo_rq_add(prod, tmp, acon + e * D);
    }
    /* line 6 */
    for (uint64_t e = 0; e < R * N; e++) {
        mul(ch->alpha, st->phi + e * D, prod);
        mul(ch->beta, phipp + e * D, tmp);
        lo_rq_add(prod, tmp, phi + e * D);
    }
    /* line 7 */
    uint32_t b[D];
    mul(ch->alpha, st->b, prod);
    mul(ch->beta, tr->b_prime_prime, tmp);
    lo_rq_add(prod, tmp, b);
    /* checks 8, 9 */
    for (uint64_t i = 0; i < R && !fc; i++)
        for (uint64_t j = 0; j < R; j++)
            if (memcmp(tr->g + (i * R + j) * D, tr->g + (j * R + i) * D, D * 4)) { fc = 8; break; }
    for (uint64_t i = 0; i < R && !fc; i++)
        for (uint64_t j = 0; j < R; j++)
            if (memcmp(tr->h + (i * R + j) * D, tr->h + (j * R + i) * D, D * 4)) { fc = 9; break; }
    /* lines 10-14: exact-integer norm of all digits (z: base B, 2 digits; t: B_1,T_1; all R^2 g: B_2,T_2; all R^2 h: B_1,T_1) */
    {
        uint64_t s = 0;
        uint32_t dig[64 * D];
        for (uint64_t n = 0; n < N; n++) { lo_decompose(tr->z + n * D, c->B, 2, dig); s += lo_norm_sq(dig, 2 * D); }
        for (uint64_t e = 0; e < R * K; e++) { lo_decompose(tr->t + e * D, c->B_1, c->T_1, dig); s += lo_norm_sq(dig, (size_t)c->T_1 * D); }
        for (uint64_t e = 0; e < R * R; e++) { lo_decompose(tr->g + e * D, c->B_2, c->T_2, dig); s += lo_norm_sq(dig, (size_t)c->T_2 * D); }
        for (uint64_t e = 0; e < R * R; e++) { lo_decompose(tr->h + e * D, c->B_1, c->T_1, dig); s += lo_norm_sq(dig, (size_t)c->T_1 * D); }
        if (norm_sum) *norm_sum = s;
        if (!fc && (double)s > c->BETA_PRIME) fc = 14;     /* :265 */
    }
    if (fc) goto done;
    /* check 15: A z == sum c_i t_i */
    {
        cand = malloc(K * D * sizeof(uint32_t));
        lo_constants c1 = *c; c1.R = 1;
        lo_commit_inner_rows(&c1, seed, tr->z, 0, K, use_ntt, nthreads, cand);
        for (uint64_t y = 0; y < K && !fc; y++) {
            uint32_t acc[D];
            memset(acc, 0, sizeof acc);
            for (uint64_t i = 0; i < R; i++) { mul(ch->c + i * D, tr->t + (i * K + y) * D, prod); lo_rq_add(acc, prod, acc); }
            if (memcmp(acc, cand + y * D, sizeof acc)) fc = 15;
        }
        free(cand); cand = NULL;
        if (fc) goto done;
    }
    /* check 16: <z,z> == sum g_ij c_i c_j */
    {
        uint32_t lhs[D], r2[D];
        ip(mul, tr->z, tr->z, N, lhs);
        memset(r2, 0, sizeof r2);
        for (uint64_t i = 0; i < R; i++)
            for (uint64_t j = 0; j < R; j++) {
                mul(tr->g + (i * R + j) * D, ch->c + i * D, prod);
                mul(prod, ch->c + j * D, tmp);
                lo_rq_add(r2, tmp, r2);
            }
        if (memcmp(lhs, r2, sizeof lhs)) { fc = 16; goto done; }
    }
    /* check 17: sum <phi_i,z> c_i == sum h_ij c_i c_j */
    {
        uint32_t lhs[D], r2[D];
        memset(lhs, 0, sizeof lhs); memset(r2, 0, sizeof r2);
        for (uint64_t i = 0; i < R; i++) {
            ip(mul, phi + i * nd, tr->z, N, tmp);
            mul(tmp, ch->c + i * D, prod);
            lo_rq_add(lhs, prod, lhs);
        }
        for (uint64_t i = 0; i < R; i++)
            for (uint64_t j = 0; j < R; j++) {
                mul(tr->h + (i * R + j) * D, ch->c + i * D, prod);
                mul(prod, ch->c + j * D, tmp);
                lo_rq_add(r2, tmp, r2);
            }
        if (memcmp(lhs, r2, sizeof lhs)) { fc = 17; goto done; }
    }
    /* check 18: sum a_ij g_ij + sum h_ii - b == 0 */
    {
        uint32_t s1[D], s2[D];
        memset(s1, 0, sizeof s1); memset(s2, 0, sizeof s2);
        for (uint64_t i = 0; i < R; i++) {
            for (uint64_t j = 0; j < R; j++) { mul(acon + (i * R + j) * D, tr->g + (i * R + j) * D, prod); lo_rq_add(s1, prod, s1); }
            lo_rq_add(s2, tr->h + (i * R + i) * D, s2);
        }
        lo_rq_add(s1, s2, s1);
        lo_rq_sub(s1, b, s1);
        for (int d = 0; d < D; d++) if (s1[d]) { fc = 18; goto done; }
    }
    /* check 19 */
    cand = malloc(K1 * D * sizeof(uint32_t));
    outer_lhs(c, seed, tr->t, mul, nthreads, cand);
    rhs = calloc(K2 * D, sizeof(uint32_t));
    outer_rank1(c, seed, tr->g, 2, c->B_2, c->T_2, mul, rhs);
    for (uint64_t x = 0; x < K1; x++) lo_rq_add(cand + x * D, rhs + x * D, cand + x * D);
    if (memcmp(cand, tr->u_1, K1 * D * 4)) { fc = 19; goto done; }
    /* check 20 */
    memset(rhs, 0, K2 * D * sizeof(uint32_t));
    outer_rank1(c, seed, tr->h, 3, c->B_1, c->T_1, mul, rhs);
    if (memcmp(rhs, tr->u_2, K2 * D * 4)) { fc = 20; goto done; }
done:
    ok = fc == 0;
    if (failed_check) *failed_check = fc;
    free(phipp); free(phi); free(acon); free(cand); free(rhs);
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * seeded input generators (distributions of SURVEY A.3; the draw order is ours)
 * stream ids: 1 witness coeffs, 2 witness reduction picks, 3 a_ij, 4 phi, 5 Pi (+attempt<<8),
 *             6 psi, 7 omega, 8 alpha, 9 beta, 10 challenge polys (+ idx<<8), 11 operator-norm samples
 * ------------------------------------------------------------------------------------------ */
/* proofgen.rs:460-518 + util.rs:27-51: uniform polys; while sum of canonical squared norms > beta^2
 * pick (n,i) uniformly and floor-halve every coefficient of that poly (Zq `/ 2` = integer division). */
void lo_generate_witness(const lo_constants *c, uint64_t seed, uint32_t *S) {
    uint64_t R = c->R, N = c->N;
    for (uint64_t e = 0; e < R * N * D; e++) S[e] = lo_prg_zq(seed, 1, e);
    __int128 norm = 0;
    for (uint64_t e = 0; e < R * N * D; e++) norm += (uint64_t)S[e] * S[e];
    __int128 bound = (__int128)c->BETA_BOUND * c->BETA_BOUND;
    uint64_t draw = 0;
    while (norm > bound) {
        uint64_t n = (uint64_t)(((u128)lo_prg_u64(seed, 2, draw++) * N) >> 64);
        uint64_t i = (uint64_t)(((u128)lo_prg_u64(seed, 2, draw++) * R) >> 64);
        uint32_t *p = S + (i * N + n) * D;
        uint64_t before = lo_norm_sq(p, D);
        for (int d = 0; d < D; d++) p[d] /= 2;
        norm -= (__int128)(before - lo_norm_sq(p, D));
    }
}
/* structs.rs:289-350: symmetric uniform a_ij, uniform phi, b = sum a_ij <s_i,s_j> + sum <phi_i,s_i> */
void lo_generate_state(const lo_constants *c, uint64_t seed, const uint32_t *S, uint32_t *phi, uint32_t *a, uint32_t *b) {
    uint64_t R = c->R, N = c->N, nd = N * D;
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t j = i; j < R; j++)
            for (int d = 0; d < D; d++) {
                uint32_t v = lo_prg_zq(seed, 3, (i * R + j) * D + (uint64_t)d);
                a[(i * R + j) * D + d] = v; a[(j * R + i) * D + d] = v;
            }
    for (uint64_t e = 0; e < R * nd; e++) phi[e] = lo_prg_zq(seed, 4, e);
    uint32_t acc[D], g[D], prod[D];
    memset(acc, 0, sizeof acc);
    for (uint64_t i = 0; i < R; i++)
        for (uint64_t j = 0; j < R; j++) {
            lo_inner_product(S + i * nd, S + j * nd, N, g);
            lo_rq_mul(a + (i * R + j) * D, g, prod);
            lo_rq_add(acc, prod, acc);
        }
    for (uint64_t i = 0; i < R; i++) { lo_inner_product(phi + i * nd, S + i * nd, N, g); lo_rq_add(acc, g, acc); }
    memcpy(b, acc, sizeof acc);
}
/* verification.rs:553-566: choices [-1,0,1] with weights [.25,.5,.25]; two PRG bits per entry:
 * 00 -> -1, 01/10 -> 0, 11 -> 1.  32 entries per PRG word, row-major fill order. */
void lo_sample_pi(const lo_constants *c, uint64_t seed, uint64_t attempt, int8_t *pi) {
    uint64_t total = c->R * LO_JL_ROWS * c->N * D;
    for (uint64_t w = 0; w * 32 < total; w++) {
        uint64_t bits = lo_prg_u64(seed, 5 + (attempt << 8), w);
        for (int t = 0; t < 32 && w * 32 + (uint64_t)t < total; t++) {
            unsigned two = (unsigned)(bits >> (2 * t)) & 3u;
            pi[w * 32 + (uint64_t)t] = (int8_t)(two == 0 ? -1 : two == 3 ? 1 : 0);
        }
    }
}
/* verification.rs:460-489 + util.rs:83-104,227-246: draw without replacement from {0 x23, 1 x31, 2 x10},
 * negate nonzero with prob 1/2; resample while the 1000-sample operator-norm estimate (ratio of
 * canonical-representative 2-norms) exceeds T = 15. */
void lo_sample_challenge_poly(uint64_t seed, uint64_t idx, uint32_t *out) {
    uint64_t draw = 0, od = 0;
    for (;;) {
        uint32_t dist[D];
        int len = 0;
        for (int t = 0; t < 23; t++) dist[len++] = 0;
        for (int t = 0; t < 31; t++) dist[len++] = 1;
        for (int t = 0; t < 10; t++) dist[len++] = 2;
        for (int d = 0; d < D; d++) {
            int ri = (int)(((u128)lo_prg_u64(seed, 10 + (idx << 8), draw++) * (uint64_t)len) >> 64);
            uint32_t coeff = dist[ri];
            memmove(dist + ri, dist + ri + 1, (size_t)(len - ri - 1) * sizeof(uint32_t));
            len--;
            int sgn = (int)(lo_prg_u64(seed, 10 + (idx << 8), draw++) >> 63);
            out[d] = (coeff > 0 && sgn) ? Q - coeff : coeff;
        }
        double sup = 0.0;
        for (int s = 0; s < 1000; s++) {
            uint32_t r[D], cr[D];
            for (int d = 0; d < D; d++) r[d] = lo_prg_zq(seed, 11 + (idx << 8), od++);
            lo_rq_mul(out, r, cr);
            double ratio = sqrt((double)lo_norm_sq(cr, D)) / sqrt((double)lo_norm_sq(r, D));
            if (ratio > sup) sup = ratio;
        }
        if (!(sup > LO_T_OPNORM)) return;
    }
}
