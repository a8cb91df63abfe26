//! Prints, as one JSON object, outputs of the UNMODIFIED reference crate that pin what this repo's CPU oracle restates from
//! third-party crates: the CRS coefficient stream (rand 0.8.5 `gen_range(0..Q)` on i128 over rand_chacha 0.3.1
//! `ChaCha20Rng::from_seed`, structs.rs:167-171), the CRS offsets (structs.rs:55-144), `RuntimeConstants::new` for every
//! `labrador_perf` shape (constants.rs:234-264; benches/labrador_perf.rs:22-28), `Rq` multiplication on both paths
//! (algebraic.rs:379-404) and the bytes of `bincode::serialize(&Transcript)` plus `size_in_bytes` (structs.rs:192-221) for a
//! fixed transcript.  tests/test_reference_vectors.py of the B200 repo diffs this output against the oracle.
use labrador_snark::algebraic::{Rq, Zq};
use labrador_snark::constants::*;
use labrador_snark::structs::{Transcript, CRS};
use ndarray::Array2;
use std::sync::atomic::Ordering;

fn coeffs(p: &Rq) -> Vec<i128> { let mut v: Vec<i128> = p.data_vec().into_iter().map(i128::from).collect(); v.resize(D as usize, 0); v }
fn json_polys(ps: &[Rq]) -> String { format!("[{}]", ps.iter().map(|p| format!("{:?}", coeffs(p))).collect::<Vec<_>>().join(",")) }
fn fnv(ps: &[Rq]) -> u64 {                      // FNV-1a 64 over the dense coefficients as u16 LE
    let mut h: u64 = 0xcbf29ce484222325;
    for p in ps { for c in coeffs(p) { for b in (c as u16).to_le_bytes() { h ^= b as u64; h = h.wrapping_mul(0x100000001b3); } } }
    h
}
fn fnv_bytes(bs: &[u8]) -> u64 { let mut h: u64 = 0xcbf29ce484222325; for &b in bs { h ^= b as u64; h = h.wrapping_mul(0x100000001b3); } h }
fn hex(bs: &[u8]) -> String { bs.iter().map(|b| format!("{:02x}", b)).collect() }

/// the fixed polynomials of the transcript vector: coefficient d of polynomial (tag, idx); trailing zeros exercise the trimming
fn poly(tag: i128, idx: i128) -> Rq {
    if idx == 3 { return Rq::new(vec![]); }
    let len = 64 - (idx % 5) as usize;
    Rq::new((0..64).map(|d| Zq::new(if d < len { tag * 131 + idx * 17 + d as i128 * 7 + 1 } else { 0 })).collect())
}

fn main() {
    let mut out: Vec<String> = vec![];
    let c22 = RuntimeConstants::new(2, 2);
    let crs = CRS::new(&c22);
    // ---- 1. the coefficient oracle, independent of the private base seed: random_oracle_gen takes the counter seed itself ----
    let mut s0 = [0u8; 32];
    let mut s1 = [0u8; 32];
    for i in 0..32 { s1[i] = i as u8; }
    let mut s2 = [0xffu8; 32];                  // + 1 carries through every byte (increment_seed, structs.rs:155-165)
    s2[0] = 0x7f;
    out.push(format!("\"random_oracle_gen\": {{\"seed_zero\": {:?}, \"seed_00_1f\": {:?}, \"seed_7fff_ff\": {:?}}}",
                     coeffs(&crs.random_oracle_gen(&mut s0)), coeffs(&crs.random_oracle_gen(&mut s1)), coeffs(&crs.random_oracle_gen(&mut s2))));
    // ---- 2. offsets: recover the private base seed from the object's bytes (the other field is the known &RuntimeConstants) ----
    let bytes: &[u8] = unsafe { std::slice::from_raw_parts(&crs as *const _ as *const u8, std::mem::size_of_val(&crs)) };
    let ptr = (&c22 as *const RuntimeConstants as usize).to_ne_bytes();
    let seed: Vec<u8> = if bytes.len() == 40 && bytes[..8] == ptr { bytes[8..].to_vec() } else if bytes.len() == 40 && bytes[32..] == ptr { bytes[..32].to_vec() }
                        else { panic!("unexpected CRS layout") };
    let (a5, b113, c011, d112) = (crs.fetch_A_row(5), crs.fetch_B_ik_row(1, 1, 3), crs.fetch_C_ijk(0, 1, 1), crs.fetch_D_ijk(1, 1, 2));
    out.push(format!("\"crs\": {{\"N\": 2, \"R\": 2, \"base_seed\": \"{}\", \"A_row_5\": {}, \"B_1_1_3_first2\": {}, \"B_1_1_3_fnv\": {}, \"C_0_1_1_first2\": {}, \"C_0_1_1_fnv\": {}, \"D_1_1_2_first2\": {}, \"D_1_1_2_fnv\": {}, \"B_0_1_0_equals_B_0_0_2\": {}}}",
                     hex(&seed), json_polys(&a5), json_polys(&b113[..2]), fnv(&b113), json_polys(&c011[..2]), fnv(&c011), json_polys(&d112[..2]), fnv(&d112),
                     coeffs(&crs.fetch_B_ik_row(0, 1, 0)[0]) == coeffs(&crs.fetch_B_ik_row(0, 0, 2)[0])));
    // ---- 3. RuntimeConstants::new for the labrador_perf shapes (size_pow 2..10) ----
    let mut cs: Vec<String> = vec![];
    let (mut n, mut r) = (1usize, 2usize);
    for pow in 2..=10 {
        if pow % 2 == 0 { n *= 2; } else { r *= 2; }
        let c = RuntimeConstants::new(n, r);
        cs.push(format!("\"{},{}\": {{\"BETA_BOUND\": {}, \"STD\": {:e}, \"B\": {}, \"T_1\": {}, \"B_1\": {}, \"T_2\": {}, \"B_2\": {}, \"GAMMA\": {:e}, \"GAMMA_1\": {:e}, \"GAMMA_2\": {:e}, \"BETA_PRIME\": {:e}, \"KAPPA\": {}}}",
                        n, r, c.BETA_BOUND, c.STD, c.B, c.T_1, c.B_1, c.T_2, c.B_2, c.GAMMA, c.GAMMA_1, c.GAMMA_2, c.BETA_PRIME, c.KAPPA));
    }
    out.push(format!("\"constants\": {{{}}}", cs.join(", ")));
    // ---- 4. Rq multiplication, schoolbook and NTT paths ----
    let (x, y) = (poly(21, 1), poly(22, 2));
    NTT_ENABLED.store(false, Ordering::SeqCst);
    let m0 = &x * &y;
    NTT_ENABLED.store(true, Ordering::SeqCst);
    let m1 = &x * &y;
    out.push(format!("\"rq_mul\": {{\"schoolbook\": {:?}, \"ntt\": {:?}}}", coeffs(&m0), coeffs(&m1)));
    // ---- 5. bincode::serialize(&Transcript) and size_in_bytes for a fixed (N, R) = (1, 2) transcript ----
    let (rr, kappa, nd) = (2usize, 64usize, 64usize);
    let polys = |tag: i128, cnt: usize| -> Vec<Rq> { (0..cnt).map(|i| poly(tag, i as i128)).collect() };
    let tr = Transcript {
        u_1: polys(1, kappa),
        pi_i_all: (0..rr).map(|i| Array2::from_shape_fn((256, nd), |(j, x)| match (i * 7 + j * 3 + x * 5) % 4 { 0 => Zq::new(-1), 3 => Zq::new(1), _ => Zq::new(0) })).collect(),
        projection: (0..256).map(|j| Zq::new(j * 37 + 5)).collect(),
        psi: vec![vec![Zq::new(123)]],
        omega: vec![(0..256).map(|j| Zq::new(j * 11 + 3)).collect()],
        b_prime_prime: polys(5, 1), alpha: polys(6, 1), beta: polys(7, 1),
        u_2: polys(8, kappa), c: polys(9, rr), z: polys(10, 1),
        t_i_all: (0..rr).map(|i| (0..kappa).map(|k| poly(11, (i * kappa + k) as i128)).collect()).collect(),
        g_mat: Array2::from_shape_fn((rr, rr), |(i, j)| poly(12, (i * rr + j) as i128)),
        h_mat: Array2::from_shape_fn((rr, rr), |(i, j)| poly(13, (i * rr + j) as i128)),
    };
    let raw = bincode::serialize(&tr).unwrap();
    out.push(format!("\"bincode\": {{\"len\": {}, \"fnv\": {}, \"head\": \"{}\", \"tail\": \"{}\", \"size_in_bytes\": {}}}",
                     raw.len(), fnv_bytes(&raw), hex(&raw[..64]), hex(&raw[raw.len() - 64..]), tr.size_in_bytes()));
    println!("{{{}}}", out.join(",\n "));
}
