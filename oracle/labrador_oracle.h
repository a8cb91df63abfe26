/*
 * labrador_oracle.h -- CPU restatement of the LaBRADOR prover hot path of
 * RatioAeterna/LaBRADOR-SNARK.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (liblabrador_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY STATUS (see DESIGN.md "Oracle"):
 *   - ring arithmetic, inner products, decomposition, sigma_{-1}, JL, every proof_gen
 *     stage and Verifier::verify: pure integer semantics restated from the cited
 *     reference lines; pinned at property level by the reference's own proptests
 *     (tests/proptest.rs:14-81), which tests/ mirrors.
 *   - ChaCha20 block function: pinned by the RFC 7539 / rand_chacha known-answer vectors.
 *   - CRS coefficient sampler (rand 0.8.5 UniformInt<i128>::sample_single over
 *     rand_chacha 0.3.1 ChaCha20Rng): the crates are NOT vendored under /root/reference
 *     and no Rust toolchain exists in the build image, and the reference holds no golden
 *     vector for it => "PARITY UNPINNED" for the CRS bit stream.  Restated from the
 *     published algorithm (documented at lo_crs_coeff).
 *
 * Dense layout everywhere: a polynomial of R_q = Z_q[X]/(X^64+1) is uint32_t[64] of
 * canonical representatives in [0,Q), Q = 8191 (reference stores a *trimmed* Vec<Zq>;
 * dense <-> trimmed conversion is a boundary concern).
 */
#ifndef LABRADOR_ORACLE_H
#define LABRADOR_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LO_D 64            /* constants.rs:15 */
#define LO_Q 8191u         /* constants.rs:195 = find_suitable_prime(2^13-1) */
#define LO_TAU 71.0        /* constants.rs:22 */
#define LO_T_OPNORM 15.0   /* constants.rs:23 */
#define LO_K 1             /* constants.rs:26 */
#define LO_L 1             /* constants.rs:32 */
#define LO_JL_ROWS 256

/* constants.rs:205-265 */
typedef struct {
    uint64_t N, R;
    int64_t BETA_BOUND;
    double STD;
    int64_t B, T_1, B_1, T_2, B_2;
    double GAMMA, GAMMA_1, GAMMA_2, BETA_PRIME;
    uint64_t KAPPA, KAPPA_1, KAPPA_2;
    int degenerate; /* 1 when the f64 formulas leave the range where the reference terminates
                       (B<2, T_1<=0, B_1<2, T_2<=0, B_2<2, or non-finite)  -- SURVEY F8 */
} lo_constants;

int lo_runtime_constants(uint64_t N, uint64_t R, lo_constants *out);

/* ---- Z_q / R_q (algebraic.rs) ---- */
uint32_t lo_mod_positive(int64_t v);                       /* util.rs:16-23 with divisor Q */
void lo_rq_mul(const uint32_t *a, const uint32_t *b, uint32_t *c);   /* algebraic.rs:379-404 (schoolbook + reduction) */
void lo_rq_add(const uint32_t *a, const uint32_t *b, uint32_t *c);   /* algebraic.rs:441-505 */
void lo_rq_sub(const uint32_t *a, const uint32_t *b, uint32_t *c);   /* algebraic.rs:507-515 */
void lo_rq_scale(const uint32_t *a, uint32_t s, uint32_t *c);        /* util.rs:176-180 */
void lo_inner_product(const uint32_t *v1, const uint32_t *v2, size_t n, uint32_t *out); /* util.rs:496-509 */
void lo_sigma_inv(const uint32_t *a, uint32_t *out);                 /* util.rs:118-137 */
/* util.rs:389-442, literal loop over the Zq operator semantics */
int  lo_decompose_literal(const uint32_t *p, int64_t base, int64_t exp, uint32_t *out /*[exp][64]*/);
/* closed form; tests assert == literal */
void lo_decompose(const uint32_t *p, int64_t base, int64_t exp, uint32_t *out /*[exp][64]*/);
uint64_t lo_norm_sq(const uint32_t *coeffs, size_t n);               /* util.rs:195-202 (canonical reps), exact */

/* fast path used only for the CPU *baseline* timing: exact negacyclic product through the
 * F_{Q^2} transform (value-identical to lo_rq_mul; plays the role of NTT_ENABLED=true) */
void lo_ntt_fwd(const uint32_t *poly, uint32_t *re_im /*[64]: re0,im0,re1,im1..*/);
void lo_ntt_inv(const uint32_t *re_im, uint32_t *poly);
void lo_rq_mul_ntt(const uint32_t *a, const uint32_t *b, uint32_t *c);
/* exponent e_j such that slot j of the transform holds f(zeta^{e_j}) */
int  lo_ntt_slot_exponent(int j);
void lo_ntt_zeta(uint32_t *re, uint32_t *im);                        /* the primitive 128th root used */

/* ---- ChaCha20 + CRS (structs.rs:27-190) ---- */
void lo_chacha20_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]);
/* coefficient at absolute offset `ctr` (128-bit) from the 256-bit big-endian base seed: structs.rs:167-171 */
uint32_t lo_crs_coeff(const uint8_t seed[32], unsigned __int128 ctr);
void lo_crs_poly(const uint8_t seed[32], unsigned __int128 start, uint32_t *out /*[64]*/);
unsigned __int128 lo_off_A(const lo_constants *c, uint64_t row);                              /* structs.rs:55-72 */
unsigned __int128 lo_off_B(const lo_constants *c, uint64_t i, uint64_t k, uint64_t row);      /* structs.rs:74-88 */
unsigned __int128 lo_off_C(const lo_constants *c, uint64_t i, uint64_t j, uint64_t k);        /* structs.rs:90-114 */
unsigned __int128 lo_off_D(const lo_constants *c, uint64_t i, uint64_t j, uint64_t k);        /* structs.rs:116-144 */
void lo_fetch_A_row(const lo_constants *c, const uint8_t seed[32], uint64_t row, uint32_t *out /*[N][64]*/);
void lo_fetch_B_ik_row(const lo_constants *c, const uint8_t seed[32], uint64_t i, uint64_t k, uint64_t row, uint32_t *out /*[KAPPA][64]*/);
void lo_fetch_C_ijk(const lo_constants *c, const uint8_t seed[32], uint64_t i, uint64_t j, uint64_t k, uint32_t *out /*[KAPPA_2][64]*/);
void lo_fetch_D_ijk(const lo_constants *c, const uint8_t seed[32], uint64_t i, uint64_t j, uint64_t k, uint32_t *out /*[KAPPA_2][64]*/);
/* helper for the python side: offsets as (lo,hi) u64 */
void lo_off_split(int which, const lo_constants *c, uint64_t i, uint64_t j, uint64_t k, uint64_t row, uint64_t *lo, uint64_t *hi);

/* ---- synthetic inputs: SplitMix64 counter PRG ---- */
#define LO_SEED 0x4C61425241444F52ull
uint64_t lo_prg_u64(uint64_t seed, uint64_t stream, uint64_t idx);
uint32_t lo_prg_zq(uint64_t seed, uint64_t stream, uint64_t idx);   /* uniform in [0,Q) (multiply-shift) */

/* ---- stage restatements (proofgen.rs) ; witness S is [R][N][64], s_i contiguous ---- */
void lo_commit_inner_rows(const lo_constants *c, const uint8_t seed[32], const uint32_t *S,
                          uint64_t row0, uint64_t nrows, int use_ntt, int nthreads,
                          uint32_t *T /* [R][nrows][64] */);                                   /* proofgen.rs:41-49 */
void lo_gram(const lo_constants *c, const uint32_t *S, uint32_t *G /*[R][R][64]*/);            /* proofgen.rs:59-70 */
void lo_jl_project(const lo_constants *c, const uint32_t *S, const int8_t *Pi /*[R][256][N*64]*/,
                   int64_t *p /*[256]*/);                                                      /* proofgen.rs:429-457 */
int  lo_valid_projection(const lo_constants *c, const int64_t *p);                             /* verification.rs:568-579 (f64, literal) */
void lo_amortize_z(const lo_constants *c, const uint32_t *S, const uint32_t *ch /*[R][64]*/, uint32_t *z /*[N][64]*/); /* proofgen.rs:380-399 */

/* statement (structs.rs:269-388) */
typedef struct {
    const uint32_t *phi;   /* [R][N][64]  phi_k[0].column(i) contiguous */
    const uint32_t *a;     /* [R][R][64]  symmetric */
    const uint32_t *b;     /* [64] */
} lo_state;

/* injected verifier randomness in consumption order */
typedef struct {
    const int8_t *pi;      /* [n_attempts][R][256][N*64] entries in {-1,0,1} */
    int n_attempts;        /* <= 6 */
    uint32_t psi;          /* L=1 */
    const uint32_t *omega; /* [256] */
    const uint32_t *alpha; /* [64] */
    const uint32_t *beta;  /* [64] */
    const uint32_t *c;     /* [R][64] */
} lo_challenges;

/* Transcript (structs.rs:192-209), dense. All buffers caller-allocated. */
typedef struct {
    uint32_t *u_1;            /* [KAPPA_1][64] */
    int      jl_attempt;      /* index of the accepted Pi attempt */
    int64_t  *projection_int; /* [256] exact integers (proofgen.rs:163) */
    uint32_t *projection;     /* [256] mod Q (proofgen.rs:186) */
    uint32_t *b_prime_prime;  /* [64] */
    uint32_t *u_2;            /* [KAPPA_2][64] */
    uint32_t *z;              /* [N][64] */
    uint32_t *t;              /* [R][KAPPA][64] */
    uint32_t *g;              /* [R][R][64] */
    uint32_t *h;              /* [R][R][64] */
    uint32_t *phi_final;      /* [R][N][64]  (not in the reference transcript; exposed for parity of G5/G6) */
} lo_transcript;

#define LO_OK 0
#define LO_ERR_JL_REJECTED 1   /* proofgen.rs:175-176 */
#define LO_ERR_BPP_CHECK 2     /* verification.rs:550 */
#define LO_ERR_PARAMS 4        /* degenerate RuntimeConstants */

int lo_prove(const lo_constants *c, const uint8_t seed[32], const uint32_t *S, const lo_state *st,
             const lo_challenges *ch, int use_ntt, int nthreads, lo_transcript *out);          /* proofgen.rs:30-427 */

/* Verifier::verify (verification.rs:25-438). Returns 1 accept / 0 reject. failed_check receives the
 * reference's check number (8..20) that rejected, or 0. norm_sum receives the exact integer of Check 14. */
int lo_verify(const lo_constants *c, const uint8_t seed[32], const lo_state *st, const lo_challenges *ch,
              const lo_transcript *tr, int use_ntt, int nthreads, int *failed_check, uint64_t *norm_sum);

/* seeded input generators restating the reference's distributions */
void lo_generate_witness(const lo_constants *c, uint64_t seed, uint32_t *S /*[R][N][64]*/);   /* proofgen.rs:460-518 */
void lo_generate_state(const lo_constants *c, uint64_t seed, const uint32_t *S,
                       uint32_t *phi, uint32_t *a, uint32_t *b);                              /* structs.rs:289-350 */
void lo_sample_pi(const lo_constants *c, uint64_t seed, uint64_t attempt, int8_t *pi /*[R][256][N*64]*/); /* verification.rs:553-566 */
void lo_sample_challenge_poly(uint64_t seed, uint64_t idx, uint32_t *c /*[64]*/);            /* verification.rs:460-489 */

int lo_num_threads_default(void);

#ifdef __cplusplus
}
#endif
#endif
