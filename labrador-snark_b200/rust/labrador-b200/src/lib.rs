//! Drop-in for the prover path of RatioAeterna/LaBRADOR-SNARK over liblabrador_b200.so (B200, CUDA).
//!
//! Same item names, argument meaning and panics as the reference's public API (file:line of each item cited):
//! `RuntimeConstants::new`, `Zq`, `Rq` with `&a * &b`, `&a + &b`, `&a - &b`, `CRS::new` / `fetch_*`, `State::new`,
//! `Verifier::new` and its challenge methods, `Prover::new(&witness, &verifier, &constants)`,
//! `Prover::proof_gen(&mut self, &State, &mut CRS) -> Transcript`, `Prover::jl_project`, `generate_witness`,
//! `polynomial_vec_inner_product`, `decompose_polynomial`, `Verifier::verify`, `Transcript::size_in_bytes`.
//! Additions the reference lacks (it is unseeded): `CRS::from_seed`, `Verifier::with_challenges`.
//! The GPU context is a thread-local default (device `LABRADOR_B200_DEVICE`, default 0): the reference API has no handle to
//! pass one through.  Never compiled in the build image (no Rust toolchain there); see INTEGRATION.md.
//!
//! Dense <-> trimmed: the reference's `Rq` is a trimmed `Vec<Zq>` (zero polynomial = empty vector, algebraic.rs:431-439).
//! `Rq::from_dense` trims trailing zeros, `Rq::to_dense` zero-pads to D = 64.
#![allow(non_snake_case)]
use labrador_b200_sys as sys;
use ndarray::Array2;
use rand::Rng;
use std::cell::RefCell;
use std::ffi::CStr;

pub const D: usize = sys::LAB_D;                                    // constants.rs:15
pub const Q: i128 = sys::LAB_Q as i128;                             // constants.rs:195
pub const JL_ROWS: usize = sys::LAB_JL_ROWS;

// ------------------------------------------------------------------ context ------------------------------------------------------------------
pub struct Context(*mut sys::lab_ctx);
impl Context {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { sys::lab_ctx_create(device, &mut p) };
        if rc != sys::LAB_OK { return Err(last_error(std::ptr::null())); }
        Ok(Context(p))
    }
    /// One process per GPU: rank 0 makes the id, the host distributes it (any transport), every rank attaches.  Afterwards
    /// `proof_gen` / `verify`, called by all ranks with the same arguments, shard rows and witness vectors inside the library.
    pub fn comm_unique_id() -> [u8; 128] {
        let mut id = [0u8; 128];
        assert_eq!(unsafe { sys::lab_comm_unique_id(id.as_mut_ptr()) }, sys::LAB_OK);
        id
    }
    pub fn comm_init(&self, id: &[u8; 128], rank: i32, world: i32) -> Result<(), String> {
        if unsafe { sys::lab_comm_init(self.0, id.as_ptr(), rank, world) } == sys::LAB_OK { Ok(()) } else { Err(last_error(self.0)) }
    }
    /// Keep transformed CRS polynomials in HBM between calls (verify right after prove, proofs under one CRS); 0 = off.
    pub fn crs_cache_configure(&self, max_bytes: usize) { unsafe { sys::lab_crs_cache_configure(self.0, max_bytes) }; }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::lab_ctx_destroy(self.0) } } }
fn last_error(ctx: *const sys::lab_ctx) -> String {
    unsafe { CStr::from_ptr(sys::lab_last_error(ctx)).to_string_lossy().into_owned() }
}
thread_local! { static CTX: RefCell<Option<Context>> = RefCell::new(None); }
/// Runs `f` with this thread's default context (created on first use; there is no CPU fallback: without a CUDA device this panics).
pub fn with_ctx<T>(f: impl FnOnce(&Context) -> T) -> T {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_none() {
            let dev = std::env::var("LABRADOR_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            *c = Some(Context::new(dev).expect("liblabrador_b200: no CUDA device"));
        }
        f(c.as_ref().unwrap())
    })
}
fn ck(ctx: &Context, rc: i32) { if rc != sys::LAB_OK { panic!("{}", last_error(ctx.0)); } }

// ------------------------------------------------------------------ constants ------------------------------------------------------------------
/// RuntimeConstants (constants.rs:205-265), the reference's field names and types.
#[derive(Clone, Debug)]
pub struct RuntimeConstants {
    pub N: usize, pub R: usize, pub BETA_BOUND: i128, pub STD: f64, pub B: i128, pub T_1: i128, pub B_1: i128, pub T_2: i128, pub B_2: i128,
    pub GAMMA: f64, pub GAMMA_1: f64, pub GAMMA_2: f64, pub BETA_PRIME: f64, pub KAPPA: usize, pub KAPPA_1: usize, pub KAPPA_2: usize,
    raw: sys::lab_constants,
}
impl RuntimeConstants {
    /// constants.rs:234.  Shapes for which the reference's formulas degenerate (B = 1: its decomposition never terminates,
    /// SURVEY F8) are returned with the same field values; the stages that need digits then refuse to run.
    pub fn new(N: usize, R: usize) -> Self {
        let mut r = sys::lab_constants::default();
        unsafe { sys::lab_runtime_constants(N as u64, R as u64, &mut r) };
        RuntimeConstants { N, R, BETA_BOUND: r.BETA_BOUND as i128, STD: r.STD, B: r.B as i128, T_1: r.T_1 as i128, B_1: r.B_1 as i128, T_2: r.T_2 as i128,
                           B_2: r.B_2 as i128, GAMMA: r.GAMMA, GAMMA_1: r.GAMMA_1, GAMMA_2: r.GAMMA_2, BETA_PRIME: r.BETA_PRIME,
                           KAPPA: r.KAPPA as usize, KAPPA_1: r.KAPPA_1 as usize, KAPPA_2: r.KAPPA_2 as usize, raw: r }
    }
}

// ------------------------------------------------------------------ Zq / Rq ------------------------------------------------------------------
/// Zq (algebraic.rs:24-54): canonical representative in an i128.
#[derive(Clone, Copy, Debug, PartialEq, Eq, PartialOrd, Ord)]
pub struct Zq { value: i128 }
impl Zq {
    pub fn new(value: i128) -> Self { Zq { value: ((value % Q) + Q) % Q } }                  // util::mod_positive (util.rs:16-23)
    pub fn zero() -> Self { Zq { value: 0 } }
    pub fn value(&self) -> i128 { self.value }
    pub fn lift(v: &Vec<i128>) -> Vec<Zq> { v.iter().map(|&x| Zq::from(x)).collect() }       // algebraic.rs:40-47
    pub fn lift_inv(v: &Vec<Zq>) -> Vec<i128> { v.iter().map(|z| z.value).collect() }        // algebraic.rs:48-54
}
impl From<i128> for Zq { fn from(v: i128) -> Self { Zq::new(v) } }
impl From<Zq> for i128 { fn from(z: Zq) -> i128 { z.value } }

/// Rq (algebraic.rs:303-376): trimmed coefficient vector of an element of Z_q[X]/(X^64+1).
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct Rq(Vec<Zq>);
impl Rq {
    /// Rq::new (algebraic.rs:308) = reduction (algebraic.rs:352-376): the coefficient of X^(kD + d) moves to X^d with sign (-1)^k.
    pub fn new(coefficients: Vec<Zq>) -> Self {
        let mut dense = [0i128; D];
        for (e, z) in coefficients.iter().enumerate() {
            let (k, d) = (e / D, e % D);
            dense[d] += if k % 2 == 0 { z.value } else { -z.value };
        }
        let mut v: Vec<Zq> = dense.iter().map(|&x| Zq::new(x)).collect();
        while matches!(v.last(), Some(z) if z.value == 0) { v.pop(); }
        Rq(v)
    }
    pub fn zero() -> Self { Rq(vec![]) }                                                     // algebraic.rs:431-439
    pub fn data_vec(&self) -> Vec<Zq> { self.0.clone() }                                     // algebraic.rs:326-328
    pub fn from_dense(c: &[u32]) -> Self {
        let mut v: Vec<Zq> = c.iter().map(|&x| Zq { value: (x % sys::LAB_Q) as i128 }).collect();
        while matches!(v.last(), Some(z) if z.value == 0) { v.pop(); }
        Rq(v)
    }
    pub fn to_dense(&self) -> [u32; D] {
        let mut out = [0u32; D];
        for (i, z) in self.0.iter().enumerate() { out[i] = z.value as u32; }
        out
    }
}
fn binop(a: &Rq, b: &Rq, f: unsafe extern "C" fn(*mut sys::lab_ctx, *const u32, *const u32, *mut u32, usize) -> i32) -> Rq {
    let (x, y, mut out) = (a.to_dense(), b.to_dense(), [0u32; D]);
    with_ctx(|ctx| ck(ctx, unsafe { f(ctx.0, x.as_ptr(), y.as_ptr(), out.as_mut_ptr(), 1) }));
    Rq::from_dense(&out)
}
/// &Rq * &Rq (algebraic.rs:517-523 -> multiply, :379-404): exact negacyclic product mod q on the GPU.  One product per call is
/// the reference's shape; anything hot should batch through `polymul_batch`.
impl std::ops::Mul for &Rq { type Output = Rq; fn mul(self, rhs: &Rq) -> Rq { binop(self, rhs, sys::lab_polymul_batch) } }
impl std::ops::Add for &Rq { type Output = Rq; fn add(self, rhs: &Rq) -> Rq { binop(self, rhs, sys::lab_rq_add_batch) } }      // algebraic.rs:441-505
impl std::ops::Sub for &Rq { type Output = Rq; fn sub(self, rhs: &Rq) -> Rq { binop(self, rhs, sys::lab_rq_sub_batch) } }      // algebraic.rs:507-515
pub fn polymul_batch(a: &[Rq], b: &[Rq]) -> Vec<Rq> {
    assert_eq!(a.len(), b.len());
    let (x, y) = (flatten(a), flatten(b));
    let mut out = vec![0u32; a.len() * D];
    with_ctx(|ctx| ck(ctx, unsafe { sys::lab_polymul_batch(ctx.0, x.as_ptr(), y.as_ptr(), out.as_mut_ptr(), a.len()) }));
    polys(&out)
}
fn flatten(v: &[Rq]) -> Vec<u32> { v.iter().flat_map(|p| p.to_dense()).collect() }
fn polys(v: &[u32]) -> Vec<Rq> { v.chunks(D).map(Rq::from_dense).collect() }
/// column-major view the C ABI wants: Array2<Rq> (N x R), column i = s_i (proofgen.rs:45) -> [R][N][64]
fn columns(m: &Array2<Rq>) -> Vec<u32> {
    let (n, r) = m.dim();
    let mut s = vec![0u32; r * n * D];
    for i in 0..r { for k in 0..n { s[(i * n + k) * D..][..D].copy_from_slice(&m[[k, i]].to_dense()); } }
    s
}
fn square(m: &Array2<Rq>) -> Vec<u32> { m.iter().flat_map(|p| p.to_dense()).collect() }      // row-major [R][R][64]

// ------------------------------------------------------------------ util.rs ------------------------------------------------------------------
/// polynomial_vec_inner_product (util.rs:496-509); panics on unequal lengths like the reference (util.rs:497-502).
pub fn polynomial_vec_inner_product(v1: &[Rq], v2: &[Rq]) -> Rq {
    assert!(v1.len() == v2.len(), "inner product not defined on vectors of unequal length. v1 length: {}, v2 length: {}", v1.len(), v2.len());
    let (x, y, mut out) = (flatten(v1), flatten(v2), [0u32; D]);
    with_ctx(|ctx| ck(ctx, unsafe { sys::lab_inner_product_batch(ctx.0, x.as_ptr(), y.as_ptr(), 1, v1.len(), out.as_mut_ptr()) }));
    Rq::from_dense(&out)
}
/// decompose_polynomial (util.rs:389-442): `exp` digit polynomials in base `base`.
pub fn decompose_polynomial(p: &Rq, base: i128, exp: i128) -> Vec<Rq> {
    let x = p.to_dense();
    let mut out = vec![0u32; exp as usize * D];
    with_ctx(|ctx| ck(ctx, unsafe { sys::lab_decompose(ctx.0, x.as_ptr(), 1, base as i64, exp as i64, out.as_mut_ptr()) }));
    polys(&out)
}

// ------------------------------------------------------------------ CRS ------------------------------------------------------------------
/// CRS (structs.rs:27-190): a 256-bit base seed; every polynomial is re-derived from the ChaCha20 counter oracle on the GPU.
pub struct CRS<'a> { base_seed: [u8; 32], constants: &'a RuntimeConstants }
impl<'a> CRS<'a> {
    pub fn new(constants: &'a RuntimeConstants) -> Self {                                    // structs.rs:173-189
        let mut base_seed = [0u8; 32];
        rand::thread_rng().fill(&mut base_seed);
        CRS { base_seed, constants }
    }
    /// Addition: the reference keeps the seed private and random; parity runs need to fix it.
    pub fn from_seed(base_seed: [u8; 32], constants: &'a RuntimeConstants) -> Self { CRS { base_seed, constants } }
    fn fetch(&self, which: u8, i: usize, j: usize, k: usize, row: usize, n: usize) -> Vec<Rq> {
        let mut out = vec![0u32; n * D];
        with_ctx(|ctx| ck(ctx, unsafe { sys::lab_crs_fetch(ctx.0, &self.constants.raw, self.base_seed.as_ptr(), which as i32, i as u64, j as u64, k as u64,
                                                           row as u64, out.as_mut_ptr()) }));
        polys(&out)
    }
    pub fn fetch_A_row(&self, row: usize) -> Vec<Rq> { self.fetch(b'A', 0, 0, 0, row, self.constants.N) }                          // structs.rs:55
    pub fn fetch_B_ik_row(&self, i: usize, k: usize, row: usize) -> Vec<Rq> { self.fetch(b'B', i, 0, k, row, self.constants.KAPPA) }  // structs.rs:74
    pub fn fetch_C_ijk(&self, i: usize, j: usize, k: usize) -> Vec<Rq> { self.fetch(b'C', i, j, k, 0, self.constants.KAPPA_2) }      // structs.rs:90
    pub fn fetch_D_ijk(&self, i: usize, j: usize, k: usize) -> Vec<Rq> { self.fetch(b'D', i, j, k, 0, self.constants.KAPPA_2) }      // structs.rs:116
}

// ------------------------------------------------------------------ State ------------------------------------------------------------------
/// State (structs.rs:269-286), K = L = 1.
pub struct State {
    pub phi_k: Vec<Array2<Rq>>, pub a_k: Vec<Array2<Rq>>, pub b_k: Vec<Rq>,
    pub phi_prime_k: Vec<Array2<Rq>>, pub a_prime_k: Vec<Array2<Rq>>, pub b_prime_k: Vec<Zq>,
}
fn uniform_poly(rng: &mut impl Rng) -> Rq { Rq::new((0..D).map(|_| Zq::new(rng.gen_range(0..Q))).collect()) }       // util.rs:27-35
impl State {
    /// State::new (structs.rs:352) -> gen_f (structs.rs:289-350): random symmetric a, random phi, b = sum a_ij <s_i, s_j> + sum <phi_i, s_i>.
    pub fn new(witness: &Array2<Rq>, constants: &RuntimeConstants) -> Self {
        let (n, r) = (constants.N, constants.R);
        let mut rng = rand::thread_rng();
        let mut a = Array2::from_elem((r, r), Rq::zero());
        for i in 0..r { for j in i..r { let p = uniform_poly(&mut rng); a[[i, j]] = p.clone(); a[[j, i]] = p; } }
        let phi = Array2::from_shape_fn((n, r), |_| uniform_poly(&mut rng));
        let s = columns(witness);
        let mut g = vec![0u32; r * r * D];
        with_ctx(|ctx| ck(ctx, unsafe { sys::lab_gram(ctx.0, &constants.raw, s.as_ptr(), g.as_mut_ptr()) }));
        let ag = polymul_batch(&a.iter().cloned().collect::<Vec<_>>(), &polys(&g));
        let mut b = [0i128; D];
        for p in &ag { for (d, z) in p.0.iter().enumerate() { b[d] += z.value; } }
        for i in 0..r {
            let (pc, sc): (Vec<Rq>, Vec<Rq>) = ((0..n).map(|k| phi[[k, i]].clone()).collect(), (0..n).map(|k| witness[[k, i]].clone()).collect());
            for (d, z) in polynomial_vec_inner_product(&pc, &sc).0.iter().enumerate() { b[d] += z.value; }
        }
        let b = Rq::new(b.iter().map(|&x| Zq::new(x)).collect());
        let b0 = b.0.first().copied().unwrap_or(Zq::zero());                                 // b.eval(0), structs.rs:373
        State { phi_k: vec![phi.clone()], a_k: vec![a.clone()], b_k: vec![b], phi_prime_k: vec![phi], a_prime_k: vec![a], b_prime_k: vec![b0] }
    }
}

// ------------------------------------------------------------------ Verifier ------------------------------------------------------------------
/// Injected verifier randomness in the order proof_gen consumes it; `None` entries fall back to thread_rng.
#[derive(Default)]
pub struct Challenges {
    pub pi: Vec<Vec<Array2<i128>>>,          // [attempt][i] -> 256 x N*64, entries in {-1, 0, 1}
    pub psi: Option<Vec<Zq>>, pub omega: Option<Vec<Zq>>, pub alpha: Option<Vec<Rq>>, pub beta: Option<Vec<Rq>>, pub c: Option<Vec<Rq>>,
}
/// Verifier (verification.rs:12-23): the prover calls it sideways for every challenge.
pub struct Verifier<'a> {
    b_prime: Option<Vec<Zq>>, constants: &'a RuntimeConstants,
    injected: RefCell<Challenges>, cursor: RefCell<(usize, usize, usize)>,       // (JL matrices handed out, challenge polys handed out, unused)
}
impl<'a> Verifier<'a> {
    pub fn new(b_prime_k: Vec<Zq>, constants: &'a RuntimeConstants) -> Self {                // verification.rs:18
        Verifier { b_prime: Some(b_prime_k), constants, injected: RefCell::new(Challenges::default()), cursor: RefCell::new((0, 0, 0)) }
    }
    /// Addition: a replayable challenge source (the reference draws everything from thread_rng).
    pub fn with_challenges(b_prime_k: Vec<Zq>, constants: &'a RuntimeConstants, ch: Challenges) -> Self {
        Verifier { b_prime: Some(b_prime_k), constants, injected: RefCell::new(ch), cursor: RefCell::new((0, 0, 0)) }
    }
    /// verification.rs:553-566: entries -1 / 0 / +1 with probability 1/4, 1/2, 1/4, filled in row-major order.
    pub fn sample_jl_projection(&self) -> Array2<i128> {
        let r = self.constants.R;
        let mut cur = self.cursor.borrow_mut();
        let (att, i) = (cur.0 / r, cur.0 % r);
        cur.0 += 1;
        if let Some(m) = self.injected.borrow().pi.get(att).and_then(|a| a.get(i)) { return m.clone(); }
        let mut rng = rand::thread_rng();
        Array2::from_shape_fn((JL_ROWS, self.constants.N * D), |_| match rng.gen_range(0..4) { 0 => -1, 3 => 1, _ => 0 })
    }
    pub fn generate_psi(&self) -> Vec<Zq> {                                                  // verification.rs:491-501 (L = 1)
        self.injected.borrow().psi.clone().unwrap_or_else(|| vec![Zq::new(rand::thread_rng().gen_range(0..Q))])
    }
    pub fn generate_omega(&self) -> Vec<Zq> {                                                // verification.rs:503-513
        self.injected.borrow().omega.clone().unwrap_or_else(|| { let mut g = rand::thread_rng(); (0..JL_ROWS).map(|_| Zq::new(g.gen_range(0..Q))).collect() })
    }
    pub fn fetch_alpha(&self) -> Vec<Rq> { self.injected.borrow().alpha.clone().unwrap_or_else(|| vec![uniform_poly(&mut rand::thread_rng())]) }   // :441
    pub fn fetch_beta(&self) -> Vec<Rq> { self.injected.borrow().beta.clone().unwrap_or_else(|| vec![uniform_poly(&mut rand::thread_rng())]) }     // :449
    /// verification.rs:460-489: 23 zeros, 31 of +-1, 10 of +-2 in random order, resampled while the 1000-sample operator-norm
    /// estimate exceeds T = 15 -- sampled on the GPU (lab_sample_challenge_polys_dev) from a fresh random seed.
    pub fn fetch_challenge(&self) -> Rq {
        let mut cur = self.cursor.borrow_mut();
        let idx = cur.1;
        cur.1 += 1;
        if let Some(c) = self.injected.borrow().c.as_ref().and_then(|v| v.get(idx)) { return c.clone(); }
        let seed: u64 = rand::thread_rng().gen();
        let mut out = [0u32; D];
        with_ctx(|ctx| unsafe {
            let mut d: *mut std::ffi::c_void = std::ptr::null_mut();
            ck(ctx, sys::lab_malloc(ctx.0, D * 4, &mut d));
            ck(ctx, sys::lab_sample_challenge_polys_dev(ctx.0, seed, idx as u32, 1, d as *mut u32, std::ptr::null_mut()));
            ck(ctx, sys::lab_memcpy_d2h(ctx.0, out.as_mut_ptr() as *mut _, d, D * 4));
            ck(ctx, sys::lab_sync(ctx.0));
            ck(ctx, sys::lab_free(ctx.0, d));
        });
        Rq::from_dense(&out)
    }
    /// verification.rs:568-579, the literal f64 test sqrt(sum p^2) <= sqrt(128) * beta.
    pub fn valid_projection(&self, projection: &Vec<i128>) -> bool {
        let ss: f64 = projection.iter().map(|&p| (p as f64) * (p as f64)).sum();
        ss.sqrt() <= (128f64).sqrt() * self.constants.BETA_BOUND as f64
    }
    /// Verifier::verify (verification.rs:25-438) on the GPU: Checks 8-20 in the reference's order.
    pub fn verify(&self, st: &State, proof: &Transcript, crs: &mut CRS) -> bool {
        let c = self.constants;
        let (n, r, nd) = (c.N, c.R, c.N * D);
        let (phi, a, b) = (columns(&st.phi_k[0]), square(&st.a_k[0]), st.b_k[0].to_dense());
        let mut pi = vec![0i8; r * JL_ROWS * nd];
        for i in 0..r { for j in 0..JL_ROWS { for x in 0..nd {
            let v = proof.pi_i_all[i][[j, x]].value;                                         // lifted: q - 1 stands for -1 (proofgen.rs:445-453)
            pi[(i * JL_ROWS + j) * nd + x] = if v == Q - 1 { -1 } else { v as i8 };
        } } }
        let omega: Vec<u32> = proof.omega[0].iter().map(|z| z.value as u32).collect();
        let (alpha, beta, cd) = (proof.alpha[0].to_dense(), proof.beta[0].to_dense(), flatten(&proof.c));
        let (mut u1, mut u2, mut z, mut g, mut h) = (flatten(&proof.u_1), flatten(&proof.u_2), flatten(&proof.z), square(&proof.g_mat), square(&proof.h_mat));
        let mut t: Vec<u32> = proof.t_i_all.iter().flat_map(|ti| flatten(ti)).collect();
        let mut pmod: Vec<u32> = proof.projection.iter().map(|z| z.value as u32).collect();
        let mut bpp = proof.b_prime_prime[0].to_dense();
        let mut pint = vec![0i64; JL_ROWS];
        let cst = sys::lab_state { phi: phi.as_ptr(), a: a.as_ptr(), b: b.as_ptr() };
        let cch = sys::lab_challenges { pi: pi.as_ptr(), n_attempts: 1, psi: proof.psi[0][0].value as u32, omega: omega.as_ptr(), alpha: alpha.as_ptr(),
                                         beta: beta.as_ptr(), c: cd.as_ptr(), pi2: std::ptr::null() };
        let tr = sys::lab_transcript { u_1: u1.as_mut_ptr(), jl_attempt: 0, projection_int: pint.as_mut_ptr(), projection: pmod.as_mut_ptr(),
            b_prime_prime: bpp.as_mut_ptr(), u_2: u2.as_mut_ptr(), z: z.as_mut_ptr(), t: t.as_mut_ptr(), g: g.as_mut_ptr(), h: h.as_mut_ptr(),
            phi_final: std::ptr::null_mut(), norm_sum: 0 };
        let (mut accepted, mut failed, mut norm) = (0i32, 0i32, 0u64);
        with_ctx(|ctx| ck(ctx, unsafe { sys::lab_verify(ctx.0, &c.raw, crs.base_seed.as_ptr(), &cst, &cch, &tr, &mut accepted, &mut failed, &mut norm) }));
        let _ = (n, self.b_prime.as_ref());
        accepted != 0
    }
}

// ------------------------------------------------------------------ Transcript ------------------------------------------------------------------
/// Transcript (structs.rs:192-209), the reference's 14 fields (+ the exact integer of Check 14).
pub struct Transcript {
    pub u_1: Vec<Rq>, pub pi_i_all: Vec<Array2<Zq>>, pub projection: Vec<Zq>, pub psi: Vec<Vec<Zq>>,
    pub omega: Vec<Vec<Zq>>, pub b_prime_prime: Vec<Rq>, pub alpha: Vec<Rq>, pub beta: Vec<Rq>, pub u_2: Vec<Rq>,
    pub c: Vec<Rq>, pub z: Vec<Rq>, pub t_i_all: Vec<Vec<Rq>>, pub g_mat: Array2<Rq>, pub h_mat: Array2<Rq>,
    pub norm_sum: u64,
    bincode: Vec<u8>, gzip_bytes: usize,
}
impl Transcript {
    /// Transcript::size_in_bytes (structs.rs:211-221): gzip(best) of bincode::serialize(&self); computed when the proof was made.
    pub fn size_in_bytes(&self) -> usize { self.gzip_bytes }
    /// the bytes bincode::serialize(&self) gives in the reference
    pub fn to_bincode(&self) -> &[u8] { &self.bincode }
}

// ------------------------------------------------------------------ Prover ------------------------------------------------------------------
/// Prover (proofgen.rs:14-28).
pub struct Prover<'a> { witness: &'a Array2<Rq>, verifier: &'a Verifier<'a>, constants: &'a RuntimeConstants }
impl<'a> Prover<'a> {
    pub fn new(witness: &'a Array2<Rq>, verifier: &'a Verifier<'a>, constants: &'a RuntimeConstants) -> Self { Prover { witness, verifier, constants } }

    /// Prover::jl_project (proofgen.rs:429-457): one attempt; (projection over Z, Pi_i lifted to Z_q).
    pub fn jl_project(&mut self) -> (Vec<i128>, Vec<Array2<Zq>>) {
        let (p, pi) = self.jl_attempt();
        let c = self.constants;
        let lifted = (0..c.R).map(|i| Array2::from_shape_fn((JL_ROWS, c.N * D), |(j, x)| Zq::new(pi[(i * JL_ROWS + j) * c.N * D + x] as i128))).collect();
        (p, lifted)
    }
    fn jl_attempt(&mut self) -> (Vec<i128>, Vec<i8>) {
        let c = self.constants;
        let nd = c.N * D;
        let mut pi = vec![0i8; c.R * JL_ROWS * nd];
        for i in 0..c.R {                                                                    // one sample_jl_projection per vector, proofgen.rs:434
            let m = self.verifier.sample_jl_projection();
            for j in 0..JL_ROWS { for x in 0..nd { pi[(i * JL_ROWS + j) * nd + x] = m[[j, x]] as i8; } }
        }
        let s = columns(self.witness);
        let mut p = vec![0i64; JL_ROWS];
        let mut acc = 0i32;
        with_ctx(|ctx| ck(ctx, unsafe { sys::lab_jl_project(ctx.0, &c.raw, s.as_ptr(), pi.as_ptr(), p.as_mut_ptr(), &mut acc) }));
        (p.iter().map(|&x| x as i128).collect(), pi)
    }

    /// Prover::proof_gen (proofgen.rs:30-427).  Panics where the reference panics: "failed JL..." after six rejected projections
    /// (proofgen.rs:175-176), the verify_b_prime_prime assert (verification.rs:550).  The verifier's randomness is drawn in the
    /// reference's order: JL attempts lazily until one is accepted (proofgen.rs:161-186), then psi, omega, alpha,
    /// beta, c_i; the accepted matrices and the other challenges then go to lab_prove in one call.
    pub fn proof_gen(&mut self, st: &State, crs: &mut CRS) -> Transcript {
        let c = self.constants;
        let (n, r, nd, kappa) = (c.N, c.R, c.N * D, c.KAPPA);
        let mut rejections = 0;
        let (_, pi) = loop {
            let (p, pi) = self.jl_attempt();
            if self.verifier.valid_projection(&p) { break (p, pi); }
            rejections += 1;
            if rejections > 5 { panic!("failed JL..."); }
        };
        let psi = self.verifier.generate_psi();
        let omega = self.verifier.generate_omega();
        let (alpha, beta) = (self.verifier.fetch_alpha(), self.verifier.fetch_beta());
        let cs: Vec<Rq> = (0..r).map(|_| self.verifier.fetch_challenge()).collect();
        let (s, phi, a, b) = (columns(self.witness), columns(&st.phi_k[0]), square(&st.a_k[0]), st.b_k[0].to_dense());
        let omega_d: Vec<u32> = omega.iter().map(|z| z.value as u32).collect();
        let (alpha_d, beta_d, cd) = (alpha[0].to_dense(), beta[0].to_dense(), flatten(&cs));
        let (mut u1, mut u2, mut z, mut t, mut g, mut h) = (vec![0u32; c.KAPPA_1 * D], vec![0u32; c.KAPPA_2 * D], vec![0u32; n * D],
            vec![0u32; r * kappa * D], vec![0u32; r * r * D], vec![0u32; r * r * D]);
        let (mut pint, mut pmod, mut bpp) = (vec![0i64; JL_ROWS], vec![0u32; JL_ROWS], vec![0u32; D]);
        let cst = sys::lab_state { phi: phi.as_ptr(), a: a.as_ptr(), b: b.as_ptr() };
        let cch = sys::lab_challenges { pi: pi.as_ptr(), n_attempts: 1, psi: psi[0].value as u32, omega: omega_d.as_ptr(),
                                         alpha: alpha_d.as_ptr(), beta: beta_d.as_ptr(), c: cd.as_ptr(), pi2: std::ptr::null() };
        let mut tr = sys::lab_transcript { u_1: u1.as_mut_ptr(), jl_attempt: 0, projection_int: pint.as_mut_ptr(), projection: pmod.as_mut_ptr(),
            b_prime_prime: bpp.as_mut_ptr(), u_2: u2.as_mut_ptr(), z: z.as_mut_ptr(), t: t.as_mut_ptr(), g: g.as_mut_ptr(), h: h.as_mut_ptr(),
            phi_final: std::ptr::null_mut(), norm_sum: 0 };
        let (mut raw, mut gz) = (Vec::new(), 0usize);
        with_ctx(|ctx| {
            let rc = unsafe { sys::lab_prove(ctx.0, &c.raw, crs.base_seed.as_ptr(), s.as_ptr(), &cst, &cch, &mut tr) };
            match rc {
                sys::LAB_OK => {}
                sys::LAB_ERR_JL_REJECTED => panic!("failed JL..."),                              // proofgen.rs:176
                sys::LAB_ERR_BPP_CHECK => panic!("verify_b_prime_prime check failed"),          // verification.rs:550
                _ => panic!("{}", last_error(ctx.0)),
            }
            let mut size = 0usize;
            assert_eq!(unsafe { sys::lab_transcript_bincode(&c.raw, &tr, &cch, std::ptr::null_mut(), 0, &mut size) }, sys::LAB_OK);
            raw = vec![0u8; size];
            assert_eq!(unsafe { sys::lab_transcript_bincode(&c.raw, &tr, &cch, raw.as_mut_ptr(), raw.len(), &mut size) }, sys::LAB_OK);
            let mut rawn = 0usize;
            assert_eq!(unsafe { sys::lab_transcript_size_in_bytes(&c.raw, &tr, &cch, &mut gz, &mut rawn) }, sys::LAB_OK);
        });
        let pi_i_all = (0..r).map(|i| Array2::from_shape_fn((JL_ROWS, nd), |(j, x)| Zq::new(pi[(i * JL_ROWS + j) * nd + x] as i128))).collect();   // proofgen.rs:445-453
        Transcript {
            u_1: polys(&u1), pi_i_all, projection: pmod.iter().map(|&x| Zq::new(x as i128)).collect(), psi: vec![psi], omega: vec![omega],
            b_prime_prime: vec![Rq::from_dense(&bpp)], alpha, beta, u_2: polys(&u2), c: cs, z: polys(&z),
            t_i_all: t.chunks(kappa * D).map(polys).collect(),
            g_mat: Array2::from_shape_vec((r, r), polys(&g)).unwrap(), h_mat: Array2::from_shape_vec((r, r), polys(&h)).unwrap(),
            norm_sum: tr.norm_sum, bincode: raw, gzip_bytes: gz,
        }
    }
}

/// generate_witness (proofgen.rs:460-518): uniform coefficients, then floor-halving of random polynomials until the sum of the
/// canonical squared norms is at most BETA_BOUND^2 -- generated on the GPU from a fresh random 64-bit seed.
pub fn generate_witness(constants: &RuntimeConstants) -> Array2<Rq> {
    let (n, r) = (constants.N, constants.R);
    let seed: u64 = rand::thread_rng().gen();
    let mut s = vec![0u32; r * n * D];
    with_ctx(|ctx| unsafe {
        let mut d: *mut std::ffi::c_void = std::ptr::null_mut();
        ck(ctx, sys::lab_malloc(ctx.0, s.len() * 4, &mut d));
        ck(ctx, sys::lab_generate_witness_dev(ctx.0, &constants.raw, seed, d as *mut u32, std::ptr::null_mut()));
        ck(ctx, sys::lab_memcpy_d2h(ctx.0, s.as_mut_ptr() as *mut _, d, s.len() * 4));
        ck(ctx, sys::lab_sync(ctx.0));
        ck(ctx, sys::lab_free(ctx.0, d));
    });
    Array2::from_shape_fn((n, r), |(k, i)| Rq::from_dense(&s[(i * n + k) * D..][..D]))
}
