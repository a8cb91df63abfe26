//! Raw FFI of include/labrador_b200.h.  Never compiled in the build image (no Rust toolchain there).
#![allow(non_camel_case_types, non_snake_case)]
use std::os::raw::{c_char, c_int, c_void};

pub const LAB_D: usize = 64;
pub const LAB_Q: u32 = 8191;
pub const LAB_JL_ROWS: usize = 256;

pub const LAB_OK: c_int = 0;
pub const LAB_ERR_JL_REJECTED: c_int = 1;
pub const LAB_ERR_BPP_CHECK: c_int = 2;
pub const LAB_ERR_SHAPE: c_int = 3;
pub const LAB_ERR_PARAMS: c_int = 4;
pub const LAB_ERR_CUDA: c_int = 5;
pub const LAB_ERR_ALLOC: c_int = 6;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct lab_constants {
    pub N: u64, pub R: u64, pub BETA_BOUND: i64, pub STD: f64,
    pub B: i64, pub T_1: i64, pub B_1: i64, pub T_2: i64, pub B_2: i64,
    pub GAMMA: f64, pub GAMMA_1: f64, pub GAMMA_2: f64, pub BETA_PRIME: f64,
    pub KAPPA: u64, pub KAPPA_1: u64, pub KAPPA_2: u64, pub degenerate: c_int,
}
#[repr(C)] pub struct lab_ctx { _private: [u8; 0] }
#[repr(C)] pub struct lab_state { pub phi: *const u32, pub a: *const u32, pub b: *const u32 }
#[repr(C)] pub struct lab_challenges {
    pub pi: *const i8, pub n_attempts: c_int, pub psi: u32,
    pub omega: *const u32, pub alpha: *const u32, pub beta: *const u32, pub c: *const u32,
}
#[repr(C)] pub struct lab_transcript {
    pub u_1: *mut u32, pub jl_attempt: c_int, pub projection_int: *mut i64, pub projection: *mut u32,
    pub b_prime_prime: *mut u32, pub u_2: *mut u32, pub z: *mut u32, pub t: *mut u32, pub g: *mut u32,
    pub h: *mut u32, pub phi_final: *mut u32, pub norm_sum: u64,
}

extern "C" {
    pub fn lab_ctx_create(device: c_int, out: *mut *mut lab_ctx) -> c_int;
    pub fn lab_ctx_destroy(ctx: *mut lab_ctx);
    pub fn lab_last_error(ctx: *const lab_ctx) -> *const c_char;
    pub fn lab_sync(ctx: *mut lab_ctx) -> c_int;
    pub fn lab_runtime_constants(N: u64, R: u64, out: *mut lab_constants) -> c_int;
    pub fn lab_ntt_fwd_batch(ctx: *mut lab_ctx, input: *const u32, out: *mut u32, n_polys: usize) -> c_int;
    pub fn lab_ntt_inv_batch(ctx: *mut lab_ctx, input: *const u32, out: *mut u32, n_polys: usize) -> c_int;
    pub fn lab_polymul_batch(ctx: *mut lab_ctx, a: *const u32, b: *const u32, c: *mut u32, n_polys: usize) -> c_int;
    pub fn lab_inner_product_batch(ctx: *mut lab_ctx, v1: *const u32, v2: *const u32, n_vecs: usize, len: usize, out: *mut u32) -> c_int;
    pub fn lab_decompose(ctx: *mut lab_ctx, input: *const u32, n_polys: usize, base: i64, exp: i64, out: *mut u32) -> c_int;
    pub fn lab_norm_sq(ctx: *mut lab_ctx, input: *const u32, n_coeffs: usize, out: *mut u64) -> c_int;
    pub fn lab_sigma_inv(ctx: *mut lab_ctx, input: *const u32, n_polys: usize, out: *mut u32) -> c_int;
    pub fn lab_crs_expand(ctx: *mut lab_ctx, seed: *const u8, start_lo: u64, start_hi: u64, n_polys: usize, out: *mut u32) -> c_int;
    pub fn lab_crs_fetch(ctx: *mut lab_ctx, c: *const lab_constants, seed: *const u8, which: c_int, i: u64, j: u64, k: u64, row: u64, out: *mut u32) -> c_int;
    pub fn lab_commit_inner(ctx: *mut lab_ctx, c: *const lab_constants, seed: *const u8, S: *const u32, row0: u64, nrows: u64, T: *mut u32) -> c_int;
    pub fn lab_gram(ctx: *mut lab_ctx, c: *const lab_constants, S: *const u32, G: *mut u32) -> c_int;
    pub fn lab_jl_project(ctx: *mut lab_ctx, c: *const lab_constants, S: *const u32, pi: *const i8, p: *mut i64, accepted: *mut c_int) -> c_int;
    pub fn lab_commit_outer_u1(ctx: *mut lab_ctx, c: *const lab_constants, seed: *const u8, T: *const u32, G: *const u32, u1: *mut u32) -> c_int;
    pub fn lab_commit_outer_u2(ctx: *mut lab_ctx, c: *const lab_constants, seed: *const u8, H: *const u32, u2: *mut u32) -> c_int;
    pub fn lab_aggregate_phi(ctx: *mut lab_ctx, c: *const lab_constants, phi: *const u32, pi: *const i8, psi: u32, omega: *const u32, phi_pp: *mut u32) -> c_int;
    pub fn lab_h_gram(ctx: *mut lab_ctx, c: *const lab_constants, phi_final: *const u32, S: *const u32, H: *mut u32) -> c_int;
    pub fn lab_amortize_z(ctx: *mut lab_ctx, c: *const lab_constants, S: *const u32, ch: *const u32, z: *mut u32) -> c_int;
    pub fn lab_prove(ctx: *mut lab_ctx, c: *const lab_constants, seed: *const u8, S: *const u32, st: *const lab_state,
                     ch: *const lab_challenges, out: *mut lab_transcript) -> c_int;
    pub fn lab_prove_batch(ctx: *mut lab_ctx, c: *const lab_constants, n_statements: usize, seeds: *const u8, shared_crs: c_int, S: *const u32,
                           st: *const lab_state, ch: *const lab_challenges, out: *mut lab_transcript) -> c_int;
    // Verifier::verify (verification.rs:25-438)
    pub fn lab_verify(ctx: *mut lab_ctx, c: *const lab_constants, seed: *const u8, st: *const lab_state, ch: *const lab_challenges,
                      tr: *const lab_transcript, accepted: *mut c_int, failed_check: *mut c_int, norm_sum: *mut u64) -> c_int;
    // one process per GPU: NCCL communicator inside the library (the host moves the 128-byte id between ranks)
    pub fn lab_comm_unique_id(id: *mut u8) -> c_int;
    pub fn lab_comm_init(ctx: *mut lab_ctx, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn lab_comm_destroy(ctx: *mut lab_ctx) -> c_int;
    // CRS cache in HBM (transparent; off unless configured)
    pub fn lab_crs_cache_configure(ctx: *mut lab_ctx, max_bytes: usize) -> c_int;
    pub fn lab_crs_cache_stats(ctx: *const lab_ctx, bytes_used: *mut usize, hits: *mut u64, misses: *mut u64) -> c_int;
    // bincode::serialize(&Transcript) byte layout (structs.rs:192-221); out = null returns the size
    pub fn lab_transcript_bincode(c: *const lab_constants, tr: *const lab_transcript, ch: *const lab_challenges,
                                  out: *mut u8, cap: usize, size: *mut usize) -> c_int;
    // seeded generators on the device: fetch_challenge (verification.rs:460-489), generate_witness (proofgen.rs:460-518), State::gen_f (structs.rs:289-350)
    pub fn lab_sample_challenge_polys_dev(ctx: *mut lab_ctx, seed: u64, first_idx: u32, count: u32, c_dev: *mut u32, candidates_dev: *mut u32) -> c_int;
    pub fn lab_generate_witness_dev(ctx: *mut lab_ctx, c: *const lab_constants, seed: u64, S_dev: *mut u32, info: *mut u64) -> c_int;
    pub fn lab_generate_state_dev(ctx: *mut lab_ctx, c: *const lab_constants, seed: u64, S_dev: *const u32, phi_dev: *mut u32, a_dev: *mut u32, b_dev: *mut u32) -> c_int;
    pub fn lab_malloc(ctx: *mut lab_ctx, bytes: usize, dptr: *mut *mut c_void) -> c_int;
    pub fn lab_free(ctx: *mut lab_ctx, dptr: *mut c_void) -> c_int;
    pub fn lab_memcpy_h2d(ctx: *mut lab_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn lab_memcpy_d2h(ctx: *mut lab_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
}
