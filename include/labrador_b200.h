/*
 * labrador_b200.h -- C ABI of the B200-native LaBRADOR prover hot path (liblabrador_b200.so).
 *
 * This is the drop-in boundary for RatioAeterna/LaBRADOR-SNARK's prover path.  The reference has
 * no FFI today (pure Rust); each entry point below names the reference item it
 * replaces (paths relative to the reference repo).  INTEGRATION.md shows the Rust `extern "C"`
 * block and the wrapper that keeps the reference's own names (Rq, CRS, Prover::proof_gen, ...).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.  Every function returns a lab_status
 *     and never unwinds.  lab_last_error(ctx) gives the message for the last failure on ctx.
 *   - a polynomial of R_q = Z_q[X]/(X^64+1), q = 8191, is uint32_t[64] of canonical
 *     representatives in [0,q) (dense form of the reference's trimmed Vec<Zq>, algebraic.rs:303-376).
 *   - witness S: uint32_t[R][N][64], s_i contiguous (reference: Array2<Rq> (N x R), column i,
 *     proofgen.rs:45).  phi: same layout.  a, g, h: uint32_t[R][R][64].  T: uint32_t[R][rows][64].
 *   - functions without a _dev suffix take HOST pointers and do their own H2D/D2H on the ctx
 *     stream and synchronise before returning.  *_dev functions take DEVICE pointers (memory
 *     from lab_malloc or any CUDA allocation on the ctx device), enqueue on the ctx stream and
 *     do NOT synchronise (call lab_sync).
 *   - a ctx is single-owner (one stream, one device, scratch memory); distinct ctxs are
 *     independent and may be used from different threads.  There is no global mutable state
 *     (the reference's NTT_ENABLED / MOD_SUSPENSION globals, constants.rs:200-201, have no
 *     equivalent: arithmetic is always the exact integer semantics).
 *   - there is no CPU fallback: without a CUDA device lab_ctx_create fails with LAB_ERR_CUDA.
 */
#ifndef LABRADOR_B200_H
#define LABRADOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAB_D 64                 /* constants.rs:15 */
#define LAB_Q 8191u              /* constants.rs:195 */
#define LAB_JL_ROWS 256          /* verification.rs:559 */

typedef enum {
    LAB_OK = 0,
    LAB_ERR_JL_REJECTED = 1,     /* panic!("failed JL...")            proofgen.rs:175-176 */
    LAB_ERR_BPP_CHECK = 2,       /* verify_b_prime_prime assert       verification.rs:550 */
    LAB_ERR_SHAPE = 3,           /* length asserts                    util.rs:262,275,299,497,512 */
    LAB_ERR_PARAMS = 4,          /* degenerate RuntimeConstants (B<2, T<=0: the reference spins or
                                    panics) or invalid arguments */
    LAB_ERR_CUDA = 5,            /* CUDA runtime failure / no device */
    LAB_ERR_ALLOC = 6
} lab_status;

/* RuntimeConstants (constants.rs:205-265), same field names */
typedef struct {
    uint64_t N, R;
    int64_t BETA_BOUND;
    double STD;
    int64_t B, T_1, B_1, T_2, B_2;
    double GAMMA, GAMMA_1, GAMMA_2, BETA_PRIME;
    uint64_t KAPPA, KAPPA_1, KAPPA_2;
    int degenerate;
} lab_constants;

typedef struct lab_ctx lab_ctx;

/* ---- context ---- */
int lab_ctx_create(int device, lab_ctx **out);
void lab_ctx_destroy(lab_ctx *ctx);
const char *lab_last_error(const lab_ctx *ctx);          /* ctx may be NULL: last create error */
int lab_sync(lab_ctx *ctx);
int lab_malloc(lab_ctx *ctx, size_t bytes, void **dptr);
int lab_free(lab_ctx *ctx, void *dptr);
int lab_memcpy_h2d(lab_ctx *ctx, void *dst, const void *src, size_t bytes);   /* async on ctx stream */
int lab_memcpy_d2h(lab_ctx *ctx, void *dst, const void *src, size_t bytes);   /* async on ctx stream */
void *lab_stream(lab_ctx *ctx);                           /* cudaStream_t, for event timing */
uint64_t lab_kernel_launches(const lab_ctx *ctx);         /* kernels launched on this ctx so far */
int lab_version(void);
/* CUDA-event stopwatch on the ctx stream (events bracket whatever is enqueued between the calls) */
int lab_timer_start(lab_ctx *ctx);
int lab_timer_stop(lab_ctx *ctx, double *elapsed_ms);     /* synchronises on the stop event */

/* ---- multi-GPU: one process per GPU, NCCL over NVLink (the reference's only parallelism is rayon, proofgen.rs:101-124) ----
 * Rank 0 calls lab_comm_unique_id and the HOST distributes the 128 bytes by its own means (MPI, a socket,
 * torch.distributed ...); every rank then attaches its ctx with lab_comm_init (collective).  From then on lab_prove and
 * lab_verify, called by all ranks with identical arguments, shard the CRS-regenerating stages by output rows --
 * rows of A for the inner commitments t_i (proofgen.rs:41-49) and for A z (verification.rs:274-279), rows of u_1
 * (proofgen.rs:101-153) and u_2 (proofgen.rs:364-378) -- and complete them with in-place all-gathers on the ctx stream;
 * every rank returns the same transcript.  Row counts the communicator size does not divide are computed unsharded.
 * libnccl.so.2 is resolved at run time; without it these three calls fail and everything else works. */
#define LAB_COMM_ID_BYTES 128
int lab_comm_unique_id(uint8_t id[LAB_COMM_ID_BYTES]);
int lab_comm_init(lab_ctx *ctx, const uint8_t id[LAB_COMM_ID_BYTES], int rank, int world);
int lab_comm_destroy(lab_ctx *ctx);
/* Collectives of the stage-sharded path, on the ctx stream, in place, device buffers (north star: "JL partial sums and
 * aggregated z are combined with NCCL over NVLink as int64 followed by mod-q reduction").  Without a communicator (or with
 * one rank) they are no-ops, so single-GPU callers run the same code.
 *   lab_comm_allreduce_i64_dev : ncclSum over int64[n]                      (JL partials: 256 words; z: N*64 words)
 *   lab_comm_allgather_dev     : rank r's bytes_per_rank bytes sit at buf + r * bytes_per_rank; all ranks get all slices
 *   lab_comm_rank              : rank / world of the attached communicator (0 / 1 without one)
 *   lab_comm_shard             : the contiguous balanced share [x0, x0 + nx) of `total` units this rank owns (the split
 *                                used by the *_sharded_dev stage calls below; no divisibility requirement) */
int lab_comm_allreduce_i64_dev(lab_ctx *ctx, int64_t *buf_dev, size_t n);
int lab_comm_allgather_dev(lab_ctx *ctx, void *buf_dev, size_t bytes_per_rank);
int lab_comm_rank(const lab_ctx *ctx, int *rank, int *world);
int lab_comm_shard(const lab_ctx *ctx, uint64_t total, uint64_t *x0, uint64_t *nx);

/* RuntimeConstants::new(N, R)  (constants.rs:234-264). Returns LAB_ERR_PARAMS (and fills out,
 * degenerate = 1) where the reference's formulas leave the range in which it terminates. */
int lab_runtime_constants(uint64_t N, uint64_t R, lab_constants *out);

/* ---- ring primitives: Rq::multiply / &Rq * &Rq (algebraic.rs:379-404, 517-523) ---- */
/* Forward transform R_q -> F_{q^2}^32.  out[p][2j], out[p][2j+1] = (re, im) of f_p(zeta^{e_j});
 * lab_ntt_slot_exponents gives e_j; zeta = 2620 + 936 i.  Replaces the per-call
 * PLAN.negacyclic_polymul forward step (algebraic.rs:396, constants.rs:197). */
int lab_ntt_fwd_batch(lab_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n_polys);
int lab_ntt_inv_batch(lab_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n_polys);
int lab_polymul_batch(lab_ctx *ctx, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t n_polys);
int lab_ntt_fwd_batch_dev(lab_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n_polys);
int lab_ntt_inv_batch_dev(lab_ctx *ctx, const uint32_t *in, uint32_t *out, size_t n_polys);
int lab_polymul_batch_dev(lab_ctx *ctx, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t n_polys);
void lab_ntt_slot_exponents(int out[32]);
/* &Rq + &Rq, &Rq - &Rq (algebraic.rs:441-515): coefficientwise mod q on dense polynomials, batched */
int lab_rq_add_batch(lab_ctx *ctx, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t n_polys);
int lab_rq_sub_batch(lab_ctx *ctx, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t n_polys);
/* polynomial_vec_inner_product (util.rs:496-509), batched: out[b] = <v1[b][0..len), v2[b][0..len)> */
int lab_inner_product_batch(lab_ctx *ctx, const uint32_t *v1, const uint32_t *v2, size_t n_vecs, size_t len, uint32_t *out);
/* decompose_polynomial (util.rs:389-442): out[k][p][64], k < exp */
int lab_decompose(lab_ctx *ctx, const uint32_t *in, size_t n_polys, int64_t base, int64_t exp, uint32_t *out);
/* poly_norm / vec_poly_norm_squared (util.rs:188-202) as an exact integer */
int lab_norm_sq(lab_ctx *ctx, const uint32_t *in, size_t n_coeffs, uint64_t *out);
int lab_norm_sq_dev(lab_ctx *ctx, const uint32_t *in, size_t n_coeffs, uint64_t *out_host);
/* sigma_inv_vec (util.rs:107-137) */
int lab_sigma_inv(lab_ctx *ctx, const uint32_t *in, size_t n_polys, uint32_t *out);

/* ---- CRS (structs.rs:27-190) ---- */
/* n_polys consecutive polynomials whose first coefficient sits at counter base_seed + start
 * (fetch_next_n / random_oracle_gen / generate_random_coeff, structs.rs:35-45,147-171) */
int lab_crs_expand(lab_ctx *ctx, const uint8_t seed[32], uint64_t start_lo, uint64_t start_hi, size_t n_polys, uint32_t *out);
int lab_crs_expand_dev(lab_ctx *ctx, const uint8_t seed[32], uint64_t start_lo, uint64_t start_hi, size_t n_polys, uint32_t *out);
/* fetch_A_row / fetch_B_ik_row / fetch_C_ijk / fetch_D_ijk (structs.rs:55-144); which = 'A','B','C','D' */
int lab_crs_fetch(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], int which,
                  uint64_t i, uint64_t j, uint64_t k, uint64_t row, uint32_t *out);
/* counter offset of the same (for callers that shard rows themselves) */
int lab_crs_offset(const lab_constants *c, int which, uint64_t i, uint64_t j, uint64_t k, uint64_t row, uint64_t *lo, uint64_t *hi);

/* ---- prover stages (proofgen.rs) ---- */
/* S1 inner Ajtai commitments t_i[row] = <A_row, s_i> for rows [row0,row0+nrows) (proofgen.rs:41-49).
 * T: [R][nrows][64].  Row-sharding across GPUs = disjoint [row0,row0+nrows) per rank.
 * Shapes from 2^22 polynomials of A (or more than 64 witness vectors) run as generate-then-contract: ChaCha20 + transform
 * into int8 limb planes, then a tcgen05 contraction, in row chunks of 4 GB; the rows of T leave for the host chunk by chunk
 * while the next chunk is generated (pinned host memory makes that copy asynchronous).  Smaller shapes: one fused kernel. */
int lab_commit_inner(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const uint32_t *S,
                     uint64_t row0, uint64_t nrows, uint32_t *T);
/* S2 garbage polynomials g_ij = <s_i, s_j>, all R^2 (proofgen.rs:59-70) */
int lab_gram(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, uint32_t *G);
/* S4 one JL attempt: p = sum_i Pi_i * coeffs(s_i) exact (proofgen.rs:429-457; util.rs:511-526).
 * pi: int8[R][256][N*64] in {-1,0,1}.  accepted receives Verifier::valid_projection (verification.rs:568-579). */
int lab_jl_project(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, const int8_t *pi, int64_t p[LAB_JL_ROWS], int *accepted);
/* the same for the witness vectors i in [i0,i0+ni) only: pi_part is int8[ni][256][N*64] (rows of Pi for those vectors),
 * p_partial the exact partial sums -- the per-rank piece of the int64 all-reduce when S4 is sharded by i */
int lab_jl_project_part(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, const int8_t *pi_part, uint64_t i0, uint64_t ni,
                        int64_t p_partial[LAB_JL_ROWS]);
/* ---- 2-bit packed JL matrices ("pi2") ----
 * Entries of Pi are in {-1,0,1} (verification.rs:553-566); 16 consecutive entries c = 16 w + k of a row share one 32-bit
 * word: bit k = (entry == +1), bit 16 + k = (entry == -1).  pi2: uint32_t[R][256][N*4] -- a quarter of the int8 bytes over
 * PCIe and HBM.  Every int8 entry point above has a packed twin; the int8 forms pack on the device and run the same kernels.
 * lab_pi_pack / lab_pi_unpack are host-side marshalling (no ctx, no arithmetic on ring elements). */
int lab_pi_pack(const int8_t *pi, size_t n_entries /* multiple of 16 */, uint32_t *pi2);
int lab_pi_unpack(const uint32_t *pi2, size_t n_entries /* multiple of 16 */, int8_t *pi);
int lab_pi_pack_dev(lab_ctx *ctx, const int8_t *pi_dev, size_t n_entries, uint32_t *pi2_dev);
int lab_jl_project2(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, const uint32_t *pi2, int64_t p[LAB_JL_ROWS], int *accepted);
int lab_jl_project2_part(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, const uint32_t *pi2_part, uint64_t i0, uint64_t ni,
                         int64_t p_partial[LAB_JL_ROWS]);
int lab_aggregate_phi2(lab_ctx *ctx, const lab_constants *c, const uint32_t *phi, const uint32_t *pi2, uint32_t psi,
                       const uint32_t omega[LAB_JL_ROWS], uint32_t *phi_pp);
/* S2 / S9 for a shard of witness vectors, host buffers: rows [i0,i0+ni) of g ([ni][R][64]) and the partial sum
 * z = sum_{i in shard} c_i s_i (canonical; partials of different shards add up mod q) */
int lab_gram_part(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, uint64_t i0, uint64_t ni, uint32_t *G_part);
int lab_amortize_z_part(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, const uint32_t *ch, uint64_t i0, uint64_t ni, uint32_t *z_partial);
/* S3 u_1 = sum B_ik dig_k(t_i) + sum_{i<=j} dig_k(g_ij) C_ijk (proofgen.rs:101-153) */
int lab_commit_outer_u1(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const uint32_t *T, const uint32_t *G, uint32_t *u1);
/* S8 u_2 = sum_{i<=j,k<T_1} dig_k(h_ij) D_ijk (proofgen.rs:364-378) */
int lab_commit_outer_u2(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const uint32_t *H, uint32_t *u2);
/* S5 phi''_i = psi phi_i + sum_j omega_j sigma_inv(pi_i^(j)) (proofgen.rs:225-256) */
int lab_aggregate_phi(lab_ctx *ctx, const lab_constants *c, const uint32_t *phi, const int8_t *pi, uint32_t psi,
                      const uint32_t omega[LAB_JL_ROWS], uint32_t *phi_pp);
/* S7 h_ij = (<phi_i,s_j> + <phi_j,s_i>) / 2 (proofgen.rs:320-358) */
int lab_h_gram(lab_ctx *ctx, const lab_constants *c, const uint32_t *phi_final, const uint32_t *S, uint32_t *H);
/* S9 z = sum_i c_i s_i (proofgen.rs:380-399) */
int lab_amortize_z(lab_ctx *ctx, const lab_constants *c, const uint32_t *S, const uint32_t *ch, uint32_t *z);

/* ---- whole proof: Prover::proof_gen (proofgen.rs:30-427) ---- */
typedef struct {                 /* State (structs.rs:269-286) with K = L = 1 */
    const uint32_t *phi;         /* [R][N][64] */
    const uint32_t *a;           /* [R][R][64] symmetric */
    const uint32_t *b;           /* [64]; b' = b[0] */
} lab_state;

/* Verifier randomness in the order proof_gen consumes it.  The reference
 * draws these from thread_rng inside Verifier (verification.rs:441-513,553-566); for bit-exact
 * parity they are injected. */
typedef struct {
    const int8_t *pi;            /* [n_attempts][R][256][N*64]; may be NULL when pi2 is given */
    int n_attempts;              /* 1..6 */
    uint32_t psi;
    const uint32_t *omega;       /* [256] */
    const uint32_t *alpha;       /* [64] */
    const uint32_t *beta;        /* [64] */
    const uint32_t *c;           /* [R][64] */
    const uint32_t *pi2;         /* optional: the same matrices 2-bit packed, [n_attempts][R][256][N*4]; used when non-NULL */
} lab_challenges;

/* Transcript (structs.rs:192-209), dense; buffers caller-allocated (host). pi_i_all is the accepted
 * attempt of lab_challenges.pi (lifted -1 -> q-1 by the wrapper), psi/omega/alpha/beta/c are the inputs. */
typedef struct {
    uint32_t *u_1;               /* [KAPPA_1][64] */
    int jl_attempt;
    int64_t *projection_int;     /* [256] */
    uint32_t *projection;        /* [256] mod q */
    uint32_t *b_prime_prime;     /* [64] */
    uint32_t *u_2;               /* [KAPPA_2][64] */
    uint32_t *z;                 /* [N][64] */
    uint32_t *t;                 /* [R][KAPPA][64] */
    uint32_t *g;                 /* [R][R][64] */
    uint32_t *h;                 /* [R][R][64] */
    uint32_t *phi_final;         /* [R][N][64] (extra, may be NULL) */
    uint64_t norm_sum;           /* exact integer of the verifier's Check 14 (verification.rs:231-267) */
} lab_transcript;

int lab_prove(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const uint32_t *S,
              const lab_state *st, const lab_challenges *ch, lab_transcript *out);

/* Batched independent proofs (BASELINE config 5): statement b uses seeds[b] (32 bytes each) or
 * seeds[0] when shared_crs != 0.  Arrays are the single-proof layouts concatenated. */
int lab_prove_batch(lab_ctx *ctx, const lab_constants *c, size_t n_statements, const uint8_t *seeds, int shared_crs,
                    const uint32_t *S, const lab_state *st, const lab_challenges *ch, lab_transcript *out);

/* Verifier::verify (verification.rs:25-438) on the GPU: recomputes lines 3-7 and runs Checks 8-20 in the reference's
 * order with early exit.  accepted = 1/0; failed_check = the reference's check number (8..20) or 0; norm_sum = the exact
 * integer of Check 14.  ch->pi must hold the attempts the transcript's jl_attempt indexes. */
int lab_verify(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const lab_state *st, const lab_challenges *ch,
               const lab_transcript *tr, int *accepted, int *failed_check, uint64_t *norm_sum);

/* ---- CRS cache: B200 has 180 GB of HBM, the reference regenerates every CRS polynomial at every use (structs.rs:35-45) ----
 * With max_bytes > 0 the outer-commitment kernel writes the transformed CRS polynomials it generates (B_ik rows, C_ijk,
 * D_ijk) through to HBM, and any later call on this ctx with the same seed, shape and row range -- Verifier::verify
 * recomputing u_1 / u_2 right after the proof (verification.rs:372-435), or further proofs under the same CRS object
 * (CRS is borrowed &mut but never mutated, proofgen.rs:30) -- streams them back at HBM speed instead of running
 * ChaCha20 (128 B per polynomial; (32,32) needs 103 GB).  The inner commitment does the same for A (as int8 limb planes)
 * and serves every later use of that A -- more than 64 witness vectors, the verifier's A z (verification.rs:274-279), the
 * next proof -- with a tcgen05 int8 contraction that streams A once from HBM (cfg 3: 34 ms instead of 3 s).  Results are
 * bit-identical.  Off by default (max_bytes = 0): a single proof then regenerates its CRS exactly like the reference.
 * Entries that do not fit stay uncached; when there is no room, entries of other seeds are dropped first.
 * Small shapes that replay lab_prove as a CUDA graph (lab_graph_stats) keep the generated CRS side of u_1 inside the graph's own
 * memory: with the cache on, a replay whose seed equals the previous replay's skips the generation (counted as a hit).
 * lab_prove_batch(shared_crs = 1) switches the cache on for its worker contexts; they keep it until a per-statement batch or
 * lab_crs_cache_configure(ctx, 0).
 * Any call also releases the few GB of transient limb planes the cold large-shape commitment keeps in the ctx
 * between calls (lab_crs_cache_configure(ctx, 0) is the "give the memory back" call). */
int lab_crs_cache_configure(lab_ctx *ctx, size_t max_bytes);
int lab_crs_cache_stats(const lab_ctx *ctx, size_t *bytes_used, uint64_t *hits, uint64_t *misses);
/* Small shapes (inputs + outputs of a proof within a few MB; (2,2) ... (8,8)) run lab_prove as ONE CUDA graph from the second
 * proof of a shape on: one H2D copy of a pinned input block, every stage of one JL attempt with the outer commitment u_1 on a
 * forked branch, one D2H copy of the output block; the CRS seed is patched into the recorded kernels per replay.  Same bits as
 * the ordinary path.  LAB_NO_GRAPH=1 disables it.  Counters (this ctx and its batch workers): graphs built, replays, and
 * whether a recording was abandoned (the ordinary path is used then). */
int lab_graph_stats(const lab_ctx *ctx, uint64_t *graphs, uint64_t *replays, int *failed);

/* ---- transcript wire format (structs.rs:192-221) ----
 * The bytes `bincode::serialize(&Transcript)` produces in the reference (bincode 1.3.3 defaults: fixed-width little-endian
 * integers, u64 sequence lengths): the 14 fields in declaration order; Rq = Vec<Zq> of the TRIMMED coefficients
 * (algebraic.rs:422-429), Zq = i128 (16 bytes); Array2<T> (ndarray 0.15 serde) = u8 version 1, two u64 dims, then the
 * elements as a row-major sequence; pi_i_all = the accepted JL attempt lifted to Z_q (-1 -> q-1, proofgen.rs:445-453);
 * t_i_all = Vec<Vec<Rq>> [R][KAPPA]; psi/omega = one inner vector each (K = L = 1).  Pure host code, no ctx.
 * Call with out == NULL to get the size.  Returns LAB_ERR_SHAPE when cap is too small. */
int lab_transcript_bincode(const lab_constants *c, const lab_transcript *tr, const lab_challenges *ch,
                           uint8_t *out, size_t cap, size_t *size);

/* Transcript::size_in_bytes (structs.rs:211-221): the bincode bytes above through gzip at best compression (main.rs:111-114
 * prints it in KB).  zlib level 9 in a gzip container; the reference's flate2 uses the miniz_oxide backend, whose output for
 * the same input may differ by a few bytes -- a size METRIC, not a byte-parity claim.  Host code, no ctx. */
int lab_transcript_size_in_bytes(const lab_constants *c, const lab_transcript *tr, const lab_challenges *ch, size_t *gzip_bytes, size_t *bincode_bytes);

/* Compact wire format of the same 14 fields: 13 bits per Z_q coefficient (little-endian bit stream), 2 bits per
 * entry of the accepted JL matrices (the packed words as they are), 3 bits per challenge coefficient when every c_i has the
 * reference's shape {0, +-1, +-2} (13 bits otherwise), the upper triangle of the symmetric g and h (an asymmetric g or h --
 * which Checks 8 / 9 would reject -- is refused with LAB_ERR_PARAMS).  32-byte header: "LB2C", version, N, R, jl_attempt, psi.
 * At (2,2): 0.08 MB against 1.2 MB of bincode.  lab_transcript_unpack fills caller-allocated buffers (projection_int is not
 * part of the reference's transcript and is left untouched; phi_final likewise).  Host code, no ctx. */
typedef struct {                 /* writable twin of lab_challenges for one (the accepted) JL attempt */
    uint32_t *pi2;               /* [R][256][N*4] */
    uint32_t psi;
    uint32_t *omega;             /* [256] */
    uint32_t *alpha, *beta;      /* [64] each */
    uint32_t *c;                 /* [R][64] */
} lab_challenges_buf;
int lab_transcript_pack(const lab_constants *c, const lab_transcript *tr, const lab_challenges *ch, uint8_t *out, size_t cap, size_t *size);
int lab_transcript_unpack(const lab_constants *c, const uint8_t *in, size_t size, lab_transcript *tr, lab_challenges_buf *ch);

/* ---- Fiat-Shamir (the reference's README lists it as TODO, README.md:12) ----
 * The verifier's randomness is derived from the transcript prefix instead of being injected: a running SHA-256 state absorbs
 * the statement and every prover message in the order of SURVEY A.1, and 64-bit seeds squeezed from it drive the seeded
 * samplers of this library (lab_synth_pi2_dev, the uniform Z_q stream, lab_sample_challenge_polys_dev):
 *     state  = SHA256("LaBRADOR-B200-FS-v1" | crs_seed | N | R | phi | a | b)                      lab_fs_init
 *     absorb : state = SHA256(state | label | data)                                                 lab_fs_absorb
 *     squeeze: seed  = first 8 bytes (LE) of SHA256(state | label | index as u32 LE)               lab_fs_squeeze
 *   absorb "u_1"; Pi of attempt t <- squeeze("pi", t) (PRG stream 5 + (0 << 8) of that seed); absorb "proj" (t, p as 256 x i64);
 *   psi, omega <- squeeze("agg", 0) (streams 6, 7); absorb "bpp"; alpha, beta <- squeeze("ab", 0) (streams 8, 9);
 *   absorb "u_2"; c_i <- squeeze("c", 0) (lab_sample_challenge_polys_dev, index i).
 * lab_prove_fs runs Prover::proof_gen against this derived verifier and also returns the challenges it derived (ch_out, accepted
 * JL attempt packed); lab_verify_fs re-derives them from the transcript and runs Verifier::verify.  Integers are little-endian,
 * polynomials dense uint32[64]. */
int lab_fs_init(const lab_constants *c, const uint8_t crs_seed[32], const lab_state *st, uint8_t state[32]);
int lab_fs_absorb(uint8_t state[32], const char *label, const void *data, size_t bytes);
int lab_fs_squeeze(const uint8_t state[32], const char *label, uint32_t index, uint64_t *seed);
int lab_prove_fs(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const uint32_t *S, const lab_state *st,
                 lab_transcript *out, lab_challenges_buf *ch_out);
int lab_verify_fs(lab_ctx *ctx, const lab_constants *c, const uint8_t seed[32], const lab_state *st, const lab_transcript *tr,
                  int *accepted, int *failed_check, uint64_t *norm_sum);

/* ---- device-resident stage API (inputs already in HBM; used for sharded / pipelined proving) ---- */
/* S_dev: uint32_t[R][N][64] on device.  Prepares the transformed witness inside ctx. */
int lab_witness_load_dev(lab_ctx *ctx, const lab_constants *c, const uint32_t *S_dev);
/* rows [row0,row0+nrows) of T for the loaded witness; T_dev: [R][nrows][64] */
int lab_commit_inner_dev(lab_ctx *ctx, const uint8_t seed[32], uint64_t row0, uint64_t nrows, uint32_t *T_dev);
/* rows i in [i0,i0+ni) of g for the loaded witness; G_dev: [ni][R][64] (the (i,j) tile sharding of S2) */
int lab_gram_dev(lab_ctx *ctx, uint64_t i0, uint64_t ni, uint32_t *G_dev);
/* i in [i0,i0+ni): partial projection of those witness vectors; p_dev: int64[256] (overwritten) */
int lab_jl_project_dev(lab_ctx *ctx, const int8_t *pi_dev, uint64_t i0, uint64_t ni, int64_t *p_dev);
/* z restricted to witness vectors [i0,i0+ni) as exact int64 partial sums are not needed: z is mod q;
 * z_dev: [N][64] canonical partial (sum over the given i range) */
int lab_amortize_z_dev(lab_ctx *ctx, const uint32_t *ch_dev, uint64_t i0, uint64_t ni, uint32_t *z_dev);
/* packed twin of lab_jl_project_dev: pi2_dev holds the rows of vectors [i0,i0+ni) only, [ni][256][N*4] */
int lab_jl_project2_dev(lab_ctx *ctx, const uint32_t *pi2_dev, uint64_t i0, uint64_t ni, int64_t *p_dev);
/* Stage calls sharded over the communicator (all ranks call them; shard = lab_comm_shard(R)); results complete on every rank:
 *   JL: pi2_part_dev = this rank's vectors; p_dev = int64 all-reduce of the partials                 (G4)
 *   z : canonical partial widened to int64, all-reduced, reduced mod q                               (G9)
 *   g : this rank's rows of the (i, j) grid, completed by an exchange over all ranks; G_dev [R][R][64] (G2)
 * Without a communicator they compute everything locally. */
int lab_jl_project_sharded_dev(lab_ctx *ctx, const uint32_t *pi2_part_dev, int64_t *p_dev);
int lab_amortize_z_sharded_dev(lab_ctx *ctx, const uint32_t *ch_dev, uint32_t *z_dev);
int lab_gram_sharded_dev(lab_ctx *ctx, uint32_t *G_dev);
/* Host-buffer twin of lab_witness_load_dev: uploads S ([R][N][64], host) into memory the ctx owns and transforms it, so
 * that the stage calls of one proof share one upload (lab_commit_inner_resident, the *_dev and *_sharded_dev calls). */
int lab_witness_load(lab_ctx *ctx, const lab_constants *c, const uint32_t *S_host);
/* lab_commit_inner for the loaded witness with a HOST destination: rows of T leave for T_host ([R][nrows][64], pinned memory
 * makes the copies asynchronous) chunk by chunk while the next chunk is generated; returns after the last copy. */
int lab_commit_inner_resident(lab_ctx *ctx, const uint8_t seed[32], uint64_t row0, uint64_t nrows, uint32_t *T_host);


/* ---- seeded synthetic inputs / device-side challenge source ----
 * SplitMix64 counter PRG, value(idx) = mix(seed + stream * 0xD1342543DE82EF95 + (idx + 1) * 0x9E3779B97F4A7C15).
 * lab_synth_zq_dev: uniform residues floor(u64 * q / 2^64) for idx in [start, start + n).
 * lab_synth_pi_dev: JL matrix entries {-1,0,1} with P = (1/4,1/2,1/4) (verification.rs:553-566), two bits per
 * entry from stream 5 + (attempt << 8), row-major fill order. */
int lab_synth_zq_dev(lab_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t start, size_t n, uint32_t *out_dev);
int lab_synth_pi_dev(lab_ctx *ctx, uint64_t seed, uint64_t attempt, uint64_t first_entry /* multiple of 32 */, size_t total, int8_t *out_dev);
/* the same entries written directly in the packed form (total a multiple of 32): out_dev uint32_t[total / 16] */
int lab_synth_pi2_dev(lab_ctx *ctx, uint64_t seed, uint64_t attempt, uint64_t first_entry /* multiple of 32 */, size_t total, uint32_t *out_dev);
/* Verifier::fetch_challenge (verification.rs:460-489) on the device, seeded: challenge polynomials first_idx ..
 * first_idx + count - 1 into c_dev [count][64]; streams 10 / 11 + (idx << 8) for the draws without replacement from
 * {0 x23, 1 x31, 2 x10} with random signs (util.rs:83-104) and for the 1000 operator-norm samples (util.rs:227-246,
 * threshold T = 15, the reference's f64 ratio).  candidates_dev (nullable): uint32[count], candidates tried. */
int lab_sample_challenge_polys_dev(lab_ctx *ctx, uint64_t seed, uint32_t first_idx, uint32_t count, uint32_t *c_dev, uint32_t *candidates_dev);
/* generate_witness (proofgen.rs:460-518) on the device, seeded: uniform coefficients (stream 1), then floor-halving of
 * polynomials picked from stream 2 until the squared norm is <= BETA_BOUND^2.  S_dev: [R][N][64].  info (nullable, host):
 * final squared norm, number of PRG draws; passing it synchronises. */
int lab_generate_witness_dev(lab_ctx *ctx, const lab_constants *c, uint64_t seed, uint32_t *S_dev, uint64_t info[2]);
/* State::gen_f (structs.rs:289-350) on the device, seeded: symmetric uniform a (stream 3), uniform phi (stream 4),
 * b = sum a_ij <s_i,s_j> + sum <phi_i,s_i>.  phi_dev [R][N][64], a_dev [R][R][64], b_dev [64]. */
int lab_generate_state_dev(lab_ctx *ctx, const lab_constants *c, uint64_t seed, const uint32_t *S_dev, uint32_t *phi_dev, uint32_t *a_dev, uint32_t *b_dev);
/* measured ALU-pipe (LOP3 + SHF) ceiling in lane-operations per second: the roofline denominator of the
 * ChaCha20-bound kernels (BASELINE.md section 2 asks for a measured INT32 figure) */
int lab_bench_alu_peak(lab_ctx *ctx, double *lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
