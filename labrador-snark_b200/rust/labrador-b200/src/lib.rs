//! Drop-in wrapper keeping the reference's names over liblabrador_b200.so.
//! Never compiled in the build image (no Rust toolchain there); see INTEGRATION.md.
//!
//! Dense <-> trimmed: the reference's `Rq` is a trimmed `Vec<Zq>` (zero polynomial = empty vector,
//! algebraic.rs:431-439).  `Rq::from_dense` trims trailing zeros, `Rq::to_dense` zero-pads to D = 64.
use labrador_b200_sys as sys;
use ndarray::Array2;
use std::ffi::CStr;

pub const D: usize = sys::LAB_D;
pub const Q: i128 = sys::LAB_Q as i128;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct Zq(pub u32);
impl Zq {
    pub fn new(v: i128) -> Self { Zq((((v % Q) + Q) % Q) as u32) }          // util::mod_positive (util.rs:16-23)
}

#[derive(Clone, Debug, PartialEq, Eq)]
pub struct Rq(pub Vec<Zq>);                                                  // trimmed, like Polynomial<Zq>
impl Rq {
    pub fn from_dense(c: &[u32]) -> Self {
        let mut v: Vec<Zq> = c.iter().map(|&x| Zq(x)).collect();
        while matches!(v.last(), Some(Zq(0))) { v.pop(); }
        Rq(v)
    }
    pub fn to_dense(&self) -> [u32; D] {
        let mut out = [0u32; D];
        for (i, z) in self.0.iter().enumerate() { out[i] = z.0; }
        out
    }
}

pub type RuntimeConstants = sys::lab_constants;
pub fn runtime_constants(n: usize, r: usize) -> RuntimeConstants {           // RuntimeConstants::new (constants.rs:234)
    let mut c = RuntimeConstants::default();
    unsafe { sys::lab_runtime_constants(n as u64, r as u64, &mut c) };
    c
}

pub struct Context(*mut sys::lab_ctx);
impl Context {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { sys::lab_ctx_create(device, &mut p) };
        if rc != sys::LAB_OK { return Err(last_error(std::ptr::null())); }
        Ok(Context(p))
    }
}
impl Context {
    /// One process per GPU: rank 0 makes the id, the host distributes it (any transport), every rank attaches.
    /// Afterwards `Prover::proof_gen` / `verify`, called by all ranks with the same arguments, shard the CRS-regenerating
    /// stages by rows inside the library and all-gather them over NVLink (the reference: rayon, proofgen.rs:101-124).
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

/// CRS (structs.rs:27-190).  `from_seed` is the addition parity needs (CRS::new draws from thread_rng).
pub struct CRS<'a> { pub base_seed: [u8; 32], pub constants: &'a RuntimeConstants }
impl<'a> CRS<'a> {
    pub fn from_seed(seed: [u8; 32], constants: &'a RuntimeConstants) -> Self { CRS { base_seed: seed, constants } }
    pub fn fetch_a_row(&self, ctx: &Context, row: usize) -> Vec<Rq> {       // fetch_A_row (structs.rs:55)
        let mut out = vec![0u32; self.constants.N as usize * D];
        let rc = unsafe { sys::lab_crs_fetch(ctx.0, self.constants, self.base_seed.as_ptr(), b'A' as i32, 0, 0, 0, row as u64, out.as_mut_ptr()) };
        assert_eq!(rc, sys::LAB_OK, "{}", last_error(ctx.0));
        out.chunks(D).map(Rq::from_dense).collect()
    }
}

/// State (structs.rs:269-286), K = L = 1.
pub struct State { pub phi_k: Vec<Array2<Rq>>, pub a_k: Vec<Array2<Rq>>, pub b_k: Vec<Rq> }

/// Verifier randomness in consumption order; the default implementation forwards to the
/// reference's Verifier RNG methods so unmodified callers keep working.
pub trait ChallengeSource {
    fn sample_jl_projection(&mut self, attempt: usize, i: usize) -> Array2<i128>;   // verification.rs:553
    fn generate_psi(&mut self) -> Vec<Zq>;                                            // verification.rs:491
    fn generate_omega(&mut self) -> Vec<Zq>;                                          // verification.rs:503
    fn fetch_alpha(&mut self) -> Vec<Rq>;                                             // verification.rs:441
    fn fetch_beta(&mut self) -> Vec<Rq>;                                              // verification.rs:449
    fn fetch_challenge(&mut self) -> Rq;                                              // verification.rs:460
}

/// Transcript (structs.rs:192-209) with the reference's field names.
pub struct Transcript {
    pub u_1: Vec<Rq>, pub pi_i_all: Vec<Array2<Zq>>, pub projection: Vec<Zq>, pub psi: Vec<Vec<Zq>>,
    pub omega: Vec<Vec<Zq>>, pub b_prime_prime: Vec<Rq>, pub alpha: Vec<Rq>, pub beta: Vec<Rq>, pub u_2: Vec<Rq>,
    pub c: Vec<Rq>, pub z: Vec<Rq>, pub t_i_all: Vec<Vec<Rq>>, pub g_mat: Array2<Rq>, pub h_mat: Array2<Rq>,
    pub norm_sum: u64,
}

pub struct Prover<'a> { pub witness: &'a Array2<Rq>, pub constants: &'a RuntimeConstants }
impl<'a> Prover<'a> {
    pub fn new(witness: &'a Array2<Rq>, constants: &'a RuntimeConstants) -> Self { Prover { witness, constants } }

    /// Prover::proof_gen (proofgen.rs:30).  Panics exactly where the reference panics.
    pub fn proof_gen(&mut self, ctx: &Context, st: &State, crs: &CRS, chal: &mut dyn ChallengeSource) -> Transcript {
        let c = self.constants;
        let (n, r) = (c.N as usize, c.R as usize);
        // witness (N x R, column i = s_i, proofgen.rs:45) -> [R][N][64]
        let mut s = vec![0u32; r * n * D];
        for i in 0..r { for k in 0..n { s[(i * n + k) * D..][..D].copy_from_slice(&self.witness[[k, i]].to_dense()); } }
        let mut phi = vec![0u32; r * n * D];
        for i in 0..r { for k in 0..n { phi[(i * n + k) * D..][..D].copy_from_slice(&st.phi_k[0][[k, i]].to_dense()); } }
        let mut a = vec![0u32; r * r * D];
        for i in 0..r { for j in 0..r { a[(i * r + j) * D..][..D].copy_from_slice(&st.a_k[0][[i, j]].to_dense()); } }
        let b = st.b_k[0].to_dense();
        // challenges, drawn in the reference's order; up to 6 JL attempts are drawn lazily by a real implementation
        let nd = n * D;
        let attempts = 6usize;
        let mut pi = vec![0i8; attempts * r * sys::LAB_JL_ROWS * nd];
        for at in 0..attempts { for i in 0..r {
            let m = chal.sample_jl_projection(at, i);
            for j in 0..sys::LAB_JL_ROWS { for x in 0..nd { pi[((at * r + i) * sys::LAB_JL_ROWS + j) * nd + x] = m[[j, x]] as i8; } }
        } }
        let psi = chal.generate_psi();
        let omega: Vec<u32> = chal.generate_omega().iter().map(|z| z.0).collect();
        let alpha = chal.fetch_alpha(); let beta = chal.fetch_beta();
        let (alpha_d, beta_d) = (alpha[0].to_dense(), beta[0].to_dense());
        let cs: Vec<Rq> = (0..r).map(|_| chal.fetch_challenge()).collect();
        let mut cd = vec![0u32; r * D];
        for i in 0..r { cd[i * D..][..D].copy_from_slice(&cs[i].to_dense()); }
        let kappa = c.KAPPA as usize;
        let (mut u1, mut u2, mut z, mut t, mut g, mut h) = (vec![0u32; kappa * D], vec![0u32; kappa * D], vec![0u32; n * D],
            vec![0u32; r * kappa * D], vec![0u32; r * r * D], vec![0u32; r * r * D]);
        let (mut pint, mut pmod, mut bpp) = (vec![0i64; 256], vec![0u32; 256], vec![0u32; D]);
        let cst = sys::lab_state { phi: phi.as_ptr(), a: a.as_ptr(), b: b.as_ptr() };
        let cch = sys::lab_challenges { pi: pi.as_ptr(), n_attempts: attempts as i32, psi: psi[0].0, omega: omega.as_ptr(),
                                         alpha: alpha_d.as_ptr(), beta: beta_d.as_ptr(), c: cd.as_ptr() };
        let mut tr = sys::lab_transcript { u_1: u1.as_mut_ptr(), jl_attempt: 0, projection_int: pint.as_mut_ptr(), projection: pmod.as_mut_ptr(),
            b_prime_prime: bpp.as_mut_ptr(), u_2: u2.as_mut_ptr(), z: z.as_mut_ptr(), t: t.as_mut_ptr(), g: g.as_mut_ptr(), h: h.as_mut_ptr(),
            phi_final: std::ptr::null_mut(), norm_sum: 0 };
        let rc = unsafe { sys::lab_prove(ctx.0, c, crs.base_seed.as_ptr(), s.as_ptr(), &cst, &cch, &mut tr) };
        match rc {
            sys::LAB_OK => {}
            sys::LAB_ERR_JL_REJECTED => panic!("failed JL..."),                                  // proofgen.rs:176
            sys::LAB_ERR_BPP_CHECK => panic!("verify_b_prime_prime check failed"),              // verification.rs:550
            _ => panic!("{}", last_error(ctx.0)),
        }
        let polys = |v: &[u32]| -> Vec<Rq> { v.chunks(D).map(Rq::from_dense).collect() };
        let at = tr.jl_attempt as usize;
        let pi_i_all = (0..r).map(|i| Array2::from_shape_fn((256, nd), |(j, x)| {
            Zq::new(pi[((at * r + i) * 256 + j) * nd + x] as i128) })).collect();                // lifted, proofgen.rs:445-453
        Transcript {
            u_1: polys(&u1), pi_i_all, projection: pmod.iter().map(|&x| Zq(x)).collect(), psi: vec![psi], omega: vec![omega.iter().map(|&x| Zq(x)).collect()],
            b_prime_prime: vec![Rq::from_dense(&bpp)], alpha, beta, u_2: polys(&u2), c: cs, z: polys(&z),
            t_i_all: t.chunks(kappa * D).map(|ti| polys(ti)).collect(),
            g_mat: Array2::from_shape_vec((r, r), polys(&g)).unwrap(), h_mat: Array2::from_shape_vec((r, r), polys(&h)).unwrap(),
            norm_sum: tr.norm_sum,
        }
    }
}

/// Verifier::verify (verification.rs:25-438) for a transcript kept in the dense ABI form: `Ok(())` or the number of the
/// reference's check that failed (8..20).  `raw` are the buffers `lab_prove` filled.
pub fn verify_dense(ctx: &Context, c: &RuntimeConstants, crs: &CRS, st: &sys::lab_state, ch: &sys::lab_challenges, raw: &sys::lab_transcript) -> Result<(), i32> {
    let (mut accepted, mut failed, mut norm) = (0i32, 0i32, 0u64);
    let rc = unsafe { sys::lab_verify(ctx.0, c, crs.base_seed.as_ptr(), st, ch, raw, &mut accepted, &mut failed, &mut norm) };
    assert_eq!(rc, sys::LAB_OK, "{}", last_error(ctx.0));
    if accepted != 0 { Ok(()) } else { Err(failed) }
}

/// The bytes `bincode::serialize(&Transcript)` gives in the reference (structs.rs:192-221), from the dense ABI form.
pub fn transcript_bincode(c: &RuntimeConstants, raw: &sys::lab_transcript, ch: &sys::lab_challenges) -> Vec<u8> {
    let mut size = 0usize;
    assert_eq!(unsafe { sys::lab_transcript_bincode(c, raw, ch, std::ptr::null_mut(), 0, &mut size) }, sys::LAB_OK);
    let mut out = vec![0u8; size];
    assert_eq!(unsafe { sys::lab_transcript_bincode(c, raw, ch, out.as_mut_ptr(), out.len(), &mut size) }, sys::LAB_OK);
    out
}
