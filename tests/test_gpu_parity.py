"""GPU parity: every C-ABI entry point of liblabrador_b200.so against the CPU oracle, bit-exact
(all arithmetic on this path is integer).  Mirrors the reference's own tests (tests/proptest.rs:14-81)
and adds value-level comparison on seeded inputs."""
import numpy as np
import pytest

import labrador_b200 as lb
from labrador_b200 import synth

pytestmark = pytest.mark.gpu
D, Q = 64, 8191
SEED32 = bytes(range(32))


def rand_polys(n, stream, seed=1234):
    return synth.prg_zq(seed, stream, n * D).reshape(n, D)


def edge_polys():
    z = np.zeros(D, np.uint32)
    one = z.copy(); one[0] = 1
    x63 = z.copy(); x63[63] = 1
    mx = np.full(D, Q - 1, np.uint32)
    alt = np.array([(Q - 1) if d % 2 else 1 for d in range(D)], np.uint32)
    return np.stack([z, one, x63, mx, alt])


# ---- ring arithmetic (algebraic.rs:379-404; proptest.rs:14-24 "NTT preserves result") ----
@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 1000, 4097])
def test_polymul_matches_schoolbook(ctx, orc, n):
    a, b = rand_polys(n, 1), rand_polys(n, 2)
    got = ctx.polymul_batch(a, b)
    ref = orc.rq_mul_batch(a, b)
    assert np.array_equal(got, ref)


def test_polymul_edge_cases(ctx, orc):
    e = edge_polys()
    a = np.repeat(e, len(e), axis=0)
    b = np.tile(e, (len(e), 1))
    assert np.array_equal(ctx.polymul_batch(a, b), orc.rq_mul_batch(a, b))
    assert ctx.polymul_batch(np.zeros((0, D), np.uint32), np.zeros((0, D), np.uint32)).shape == (0, D)


def test_noncanonical_inputs_are_reduced(ctx, orc):
    """Inputs are documented canonical, but any u32 is still well defined (its value mod q): lanes that see a value of
    more than 14 bits reduce first.  Mixed batches exercise both sides of that branch within one warp."""
    rng = np.random.default_rng(7)
    a = rng.integers(0, 2**32, size=(300, D), dtype=np.uint64).astype(np.uint32)
    a[::3] = rand_polys(100, 31)                      # every third polynomial canonical
    a[1, 5] = 8191; a[4, 0] = 16383; a[7, 63] = 16384
    b = rand_polys(300, 32)
    ar = (a.astype(np.uint64) % Q).astype(np.uint32)
    assert np.array_equal(ctx.polymul_batch(a, b), orc.rq_mul_batch(ar, b))
    assert np.array_equal(ctx.ntt_fwd_batch(a), np.stack([orc.ntt_fwd(p) for p in ar]))
    f = ctx.ntt_fwd_batch(ar)
    g = f + np.uint32(Q) * rng.integers(0, 4, size=f.shape, dtype=np.uint64).astype(np.uint32)   # same residues, non-canonical
    assert np.array_equal(ctx.ntt_inv_batch(g), ar)


def test_ntt_matches_oracle_and_roundtrips(ctx, orc):
    a = np.concatenate([rand_polys(300, 3), edge_polys()])
    f = ctx.ntt_fwd_batch(a)
    ref = np.stack([orc.ntt_fwd(p) for p in a])
    assert np.array_equal(f, ref)
    assert np.array_equal(ctx.ntt_inv_batch(f), a)
    # slot j is the evaluation at zeta^e_j: check slot 0 by direct evaluation in F_{Q^2}
    exps = orc.ntt_slot_exponents()
    assert exps[0] == 1


def test_inner_product_linearity(ctx, orc):
    """proptest.rs:37-64: <a, c*b> == c*<a,b> for 16-long vectors, plus value parity."""
    a, b = rand_polys(16, 4).reshape(1, 16, D), rand_polys(16, 5).reshape(1, 16, D)
    c = 4321
    ab = ctx.inner_product_batch(a, b)[0]
    assert np.array_equal(ab, orc.inner_product(a[0], b[0]))
    cb = ((b.astype(np.uint64) * c) % Q).astype(np.uint32)
    lhs = ctx.inner_product_batch(a, cb)[0]
    assert np.array_equal(lhs, ((ab.astype(np.uint64) * c) % Q).astype(np.uint32))
    with pytest.raises(lb.LabError):
        ctx.inner_product_batch(a, b[:, :15])        # util.rs:497 assert -> LAB_ERR_SHAPE


def test_sigma_inv_invariant(ctx, orc):
    """proptest.rs:68-81: <a,b>_{Z_q} == const term of <sigma_inv(a), b>_{R_q}."""
    a, b = rand_polys(16, 6), rand_polys(16, 7)
    sa = ctx.sigma_inv(a)
    assert np.array_equal(sa, np.stack([orc.sigma_inv(p) for p in a]))
    ip = ctx.inner_product_batch(sa.reshape(1, 16, D), b.reshape(1, 16, D))[0]
    assert int(ip[0]) == int((a.astype(np.uint64) * b).sum() % Q)


@pytest.mark.parametrize("base,exp", [(9, 4), (14, 2), (2, 13), (4, 6), (173, 2), (3, 2)])
def test_decompose(ctx, orc, base, exp):
    p = np.concatenate([rand_polys(20, 8), edge_polys()])
    got = ctx.decompose(p, base, exp)
    for k in range(len(p)):
        ref = orc.decompose(p[k], base, exp, literal=True)
        assert np.array_equal(got[:, k, :], ref)
    with pytest.raises(lb.LabError):
        ctx.decompose(p, 1, 4)                       # b = 1 never terminates in the reference


def test_norm_sq(ctx, orc):
    x = rand_polys(1000, 9)
    assert ctx.norm_sq(x) == orc.norm_sq(x)
    assert ctx.norm_sq(np.zeros(0, np.uint32)) == 0
    assert ctx.norm_sq(np.full(1 << 20, Q - 1, np.uint32)) == (1 << 20) * (Q - 1) ** 2


# ---- CRS (structs.rs:35-171) ----
@pytest.mark.parametrize("start", [0, 1, 63, 2**32 - 3, 2**64 - 70, 2**64 + 5, 2**100 + 12345])
def test_crs_expand(ctx, orc, start):
    got = ctx.crs_expand(SEED32, start, 5)
    assert np.array_equal(got, orc.crs_polys(SEED32, start, 5))


def test_crs_expand_carry_into_high_limbs(ctx, orc):
    seed = bytes([0]) * 8 + bytes([0xFF]) * 24      # +1 carries through three limbs
    got = ctx.crs_expand(seed, 0, 2)
    assert np.array_equal(got, orc.crs_polys(seed, 0, 2))


def test_crs_rejection_path(ctx, orc):
    """~2^-13 of the coefficients need a second 128-bit draw; 2^17 coefficients exercise it."""
    got = ctx.crs_expand(SEED32, 7, 2048)
    assert np.array_equal(got, orc.crs_polys(SEED32, 7, 2048))


def test_crs_draws_the_top_keystream_word_does_not_decide(ctx, orc):
    """The device samples from keystream word 3 alone (lab_sample_w3) and sends a coefficient to the generic path when
    L = lo32(w3 * Q) >= 0xFFF80000 - (Q - 1): there words 0..2 decide between accept (most of the band) and reject.  2^22
    coefficients, classified with the independent numpy ChaCha20 of oracle/pyref.py, contain both kinds; all of them must equal
    the oracle's literal sample_single."""
    import pyref
    n_polys, start = 1 << 16, 11
    n = n_polys * D
    low = int.from_bytes(SEED32[24:], "big") + start
    ctr = np.uint64(low) + np.arange(n, dtype=np.uint64)
    keys = np.empty((n, 8), np.uint32)
    keys[:, :6] = np.frombuffer(SEED32[:24], "<u4")
    keys[:, 6] = (ctr >> np.uint64(32)).astype(np.uint32).byteswap()
    keys[:, 7] = (ctr & np.uint64(0xFFFFFFFF)).astype(np.uint32).byteswap()
    w = pyref.chacha20_blocks(keys)[:, :4].astype(np.uint64)
    L = (w[:, 3] * np.uint64(Q)) & np.uint64(0xFFFFFFFF)
    band = np.nonzero(L >= np.uint64(0xFFF80000 - (Q - 1)))[0]
    zone = (Q << 115) - 1
    rejected = 0
    for t in band:
        v = int(w[t, 0]) | int(w[t, 1]) << 32 | int(w[t, 2]) << 64 | int(w[t, 3]) << 96
        rejected += ((v * Q) & ((1 << 128) - 1)) > zone
    assert rejected >= 100 and len(band) - rejected >= 1, (len(band), rejected)
    ref = orc.crs_polys(SEED32, start, n_polys)
    fast = np.ones(n, bool); fast[band] = False
    assert np.array_equal(ref.reshape(-1)[fast], ((w[:, 3] * np.uint64(Q)) >> np.uint64(32)).astype(np.uint32)[fast])
    got = ctx.crs_expand(SEED32, start, n_polys)
    assert np.array_equal(got, ref)


def test_crs_fetch_offsets(ctx, orc):
    c = lb.RuntimeConstants.new(2, 3)
    co, _ = orc.constants(2, 3)
    crs = lb.CRS.from_seed(c, SEED32, ctx)
    assert np.array_equal(crs.fetch_A_row(5), orc.fetch_A_row(co, SEED32, 5))
    assert np.array_equal(crs.fetch_B_ik_row(2, 1, 7), orc.fetch_B_ik_row(co, SEED32, 2, 1, 7))
    assert np.array_equal(crs.fetch_C_ijk(1, 2, 1), orc.fetch_C_ijk(co, SEED32, 1, 2, 1))
    assert np.array_equal(crs.fetch_D_ijk(0, 2, 3), orc.fetch_D_ijk(co, SEED32, 0, 2, 3))


def test_crs_32bit_boundary_of_seed_plus_counter(ctx, orc):
    """The trimmed ChaCha20 path hoists everything that depends only on key words 0..6; key word 6 changes when the low
    32 bits of seed + counter wrap.  Seeds chosen so that this happens (without a 64-bit carry) inside polynomial 5 of A,
    inside a B_ik row and inside a plain expansion: lanes before the wrap use the hoisted state, lanes after it the
    generic path, the next polynomial a recomputed hoist."""
    N, R = 4, 2
    low = (5 << 32) + (1 << 32) - 64 * 5 - 20
    seed = bytes(range(7, 31)) + low.to_bytes(8, "big")
    c = lb.RuntimeConstants.new(N, R)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=9)
    assert np.array_equal(ctx.commit_inner(c, seed, S, 0, 8), orc.commit_inner_rows(co, seed, S, 0, 8))
    assert np.array_equal(ctx.crs_expand(seed, 64 * 3, 6), orc.crs_polys(seed, 64 * 3, 6))
    # u_1 walks B_ik / C_ijk far away from counter 0: put the wrap inside row 0 of B_00
    crs = lb.CRS.from_seed(c, seed, ctx)
    assert np.array_equal(crs.fetch_B_ik_row(0, 0, 0), orc.fetch_B_ik_row(co, seed, 0, 0, 0))
    # u_1 for a seed whose wrap falls into the B region: low32(seed) + kappa*N*64 + 700 == 2^32
    startB = c.KAPPA * N * 64
    low2 = (9 << 32) + (1 << 32) - (startB % (1 << 32)) - 700
    seed2 = bytes(range(3, 27)) + low2.to_bytes(8, "big")
    Sw = orc.generate_witness(co, 9)
    st_phi, st_a, st_b = orc.generate_state(co, Sw, 3)
    ch = orc.sample_challenges(co, 11, 3)
    rc, ref = orc.prove(co, seed2, Sw, st_phi, st_a, st_b, ch, ntt=True, nthreads=8)
    assert rc == 0
    assert np.array_equal(ctx.commit_outer_u1(c, seed2, ref["t"], ref["g"]), ref["u_1"])


def test_crs_expand_hoist_cache_across_grid_stride_iterations(ctx, orc):
    """2.6 M coefficients = more than one grid-stride sweep of k_crs_expand, starting 1.5 M before a 2^32 boundary of
    seed + counter: threads refresh their hoisted state between iterations."""
    low = (3 << 32) + (1 << 32) - 1_500_000
    seed = bytes(range(11, 35)) + low.to_bytes(8, "big")
    n = 40_000
    got = ctx.crs_expand(seed, 0, n)
    ref = orc.crs_polys(seed, 0, n)
    assert np.array_equal(got, ref)


# ---- stages ----
@pytest.mark.parametrize("N,R", [(1, 1), (2, 2), (3, 5), (2, 9), (5, 17), (2, 40)])
def test_commit_inner(ctx, orc, N, R):
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=N * 100 + R)
    nrows = min(c.KAPPA, 37)
    got = ctx.commit_inner(c, SEED32, S, row0=3, nrows=nrows - 3)
    ref = orc.commit_inner_rows(co, SEED32, S, 3, nrows - 3, nthreads=8)
    assert np.array_equal(got, ref)


def test_commit_inner_counter_carry_inside_a_polynomial(ctx, orc):
    """Seed whose low 64-bit limb overflows in the middle of polynomial 3 (lanes >= 10): the per-lane offset addition
    carries into the upper key words and must take the generic path."""
    N, R = 2, 2
    low = (1 << 64) - 64 * 3 - 10
    seed = bytes(range(1, 25)) + low.to_bytes(8, "big")
    c = lb.RuntimeConstants.new(N, R)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=5)
    assert np.array_equal(ctx.commit_inner(c, seed, S, 0, 8), orc.commit_inner_rows(co, seed, S, 0, 8))
    assert np.array_equal(ctx.crs_expand(seed, 64 * 2, 4), orc.crs_polys(seed, 64 * 2, 4))
    crs = lb.CRS.from_seed(c, seed, ctx)
    assert np.array_equal(crs.fetch_B_ik_row(0, 0, 0), orc.fetch_B_ik_row(co, seed, 0, 0, 0))


@pytest.mark.parametrize("N,R", [(1, 1), (2, 2), (7, 3), (33, 5)])
def test_gram_z_jl(ctx, orc, N, R):
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=77)
    assert np.array_equal(ctx.gram(c, S), orc.gram(co, S))
    ch = rand_polys(R, 11)
    assert np.array_equal(ctx.amortize_z(c, S, ch), orc.amortize_z(co, S, ch))
    pi = synth.sample_pi(N, R, seed=5)
    p, acc = ctx.jl_project(c, S, pi)
    assert np.array_equal(p, orc.jl_project(co, S, pi))
    assert acc == orc.valid_projection(co, p)
    # sharded by witness vector: the partial sums add up exactly (the multi-GPU combine)
    h = R // 2
    parts = ctx.jl_project_part(c, S, pi[:h], 0, h) + ctx.jl_project_part(c, S, pi[h:], h, R - h)
    assert np.array_equal(parts, p)


def full_case(orc, N, R, seed, n_attempts=2):
    co, rc = orc.constants(N, R)
    assert rc == 0
    S = orc.generate_witness(co, seed)
    phi, a, b = orc.generate_state(co, S, seed)
    ch = orc.sample_challenges(co, seed, n_attempts)
    return co, S, phi, a, b, ch


@pytest.mark.parametrize("N,R", [(1, 1), (1, 2), (2, 2), (2, 3), (4, 4)])
def test_full_proof_matches_oracle_and_verifies(ctx, orc, N, R):
    """Prover::proof_gen on the GPU == the restated reference prover on the same injected inputs, and the
    restated Verifier::verify accepts the GPU transcript (main.rs:106-107)."""
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=1000 + 10 * N + R)
    c = lb.RuntimeConstants.new(N, R)
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    crs = lb.CRS.from_seed(c, SEED32, ctx)
    st = lb.State(phi, a, b)
    ver = lb.Verifier.new(st.b_prime_k, c, challenges=ch)
    tr = lb.Prover.new(S, ver, c, ctx).proof_gen(st, crs)
    got = tr.as_oracle_dict()
    for k in ("t", "g", "u_1", "projection_int", "projection", "b_prime_prime", "phi_final", "h", "u_2", "z"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["jl_attempt"] == ref["jl_attempt"]
    ok, failed, norm_sum = orc.verify(co, SEED32, phi, a, b, ch, got, ntt=True, nthreads=8)
    assert ok and failed == 0
    assert tr.norm_sum == norm_sum                 # exact-integer Check 14
    assert float(norm_sum) <= c.BETA_PRIME


def test_stage_entry_points(ctx, orc):
    """The per-stage entry points reproduce the corresponding transcript fields."""
    N, R = 2, 2
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=4242)
    c = lb.RuntimeConstants.new(N, R)
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    assert np.array_equal(ctx.commit_outer_u1(c, SEED32, ref["t"], ref["g"]), ref["u_1"])
    assert np.array_equal(ctx.commit_outer_u2(c, SEED32, ref["h"]), ref["u_2"])
    assert np.array_equal(ctx.h_gram(c, ref["phi_final"], S), ref["h"])
    pp = ctx.aggregate_phi(c, phi, ch["pi"][ref["jl_attempt"]], ch["psi"], ch["omega"])
    # phi_final = alpha*phi + beta*phi'' (proofgen.rs:301-314)
    al = np.tile(ch["alpha"], (R * N, 1)); be = np.tile(ch["beta"], (R * N, 1))
    pf = (ctx.polymul_batch(al, phi.reshape(-1, D)).astype(np.uint64) + ctx.polymul_batch(be, pp.reshape(-1, D))) % Q
    assert np.array_equal(pf.astype(np.uint32).reshape(R, N, D), ref["phi_final"])


def test_jl_rejection_paths(ctx, orc):
    """A uniform (non-short) witness is rejected by valid_projection: six rejections -> LAB_ERR_JL_REJECTED
    (proofgen.rs:169-181)."""
    N, R = 1, 2
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=99, n_attempts=6)
    S_big = synth.uniform_witness(N, R, seed=3)
    phi2, a2, b2 = orc.generate_state(co, S_big, 99)
    c = lb.RuntimeConstants.new(N, R)
    st = lb.State(phi2, a2, b2)
    ver = lb.Verifier.new(st.b_prime_k, c, challenges=ch)
    with pytest.raises(lb.LabError) as e:
        lb.Prover.new(S_big, ver, c, ctx).proof_gen(st, lb.CRS.from_seed(c, SEED32, ctx))
    assert e.value.status == 1
    rc, _ = orc.prove(co, SEED32, S_big, phi2, a2, b2, ch)
    assert rc == 1


def test_degenerate_constants_are_refused(ctx):
    c = lb.RuntimeConstants.new(4096, 64, allow_degenerate=True)
    assert c.degenerate == 1 and c.B == 1
    with pytest.raises(lb.LabError):
        lb.RuntimeConstants.new(4096, 64)


def test_gpu_verifier_matches_oracle_verifier(ctx, orc):
    """lab_verify (Verifier::verify, verification.rs:25-438) accepts honest transcripts and rejects tampered ones at the
    same check number as the restated reference verifier."""
    N, R = 2, 2
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=777)
    c = lb.RuntimeConstants.new(N, R)
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    ok, fc, ns = ctx.verify(c, SEED32, phi, a, b, ch, ref)
    ook, ofc, ons = orc.verify(co, SEED32, phi, a, b, ch, ref, ntt=True, nthreads=8)
    assert (ok, fc, ns) == (ook, ofc, ons) == (True, 0, ons)
    for field, idx in (("g", (0, 1, 3)), ("h", (1, 0, 5)), ("z", (1, 7)), ("t", (1, 17, 2)), ("g", (0, 0, 0)), ("h", (1, 1, 9)),
                       ("u_1", (5, 5)), ("u_2", (127, 63)), ("b_prime_prime", (3,))):
        bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ref.items()}
        bad[field][idx] = (int(bad[field][idx]) + 1) % Q
        if field in ("g", "h") and idx[0] != idx[1]:
            pass                                        # asymmetric tamper -> check 8 / 9
        elif field in ("g", "h"):
            pass                                        # symmetric entry: caught by a later algebraic check
        got = ctx.verify(c, SEED32, phi, a, b, ch, bad)
        want = orc.verify(co, SEED32, phi, a, b, ch, bad, ntt=True, nthreads=8)
        assert got[0] is False and got[:2] == want[:2], (field, got, want)


def test_prove_batch(ctx, orc):
    """lab_prove_batch (BASELINE config 5 shape): independent statements, per-statement and shared CRS seeds."""
    N, R, B = 1, 2, 3
    co, _ = orc.constants(N, R)
    c = lb.RuntimeConstants.new(N, R)
    cases = [full_case(orc, N, R, seed=300 + i) for i in range(B)]
    S = np.stack([cs[1] for cs in cases]); phi = np.stack([cs[2] for cs in cases])
    a = np.stack([cs[3] for cs in cases]); b = np.stack([cs[4] for cs in cases])
    chs = [cs[5] for cs in cases]
    seeds = [bytes([i]) * 32 for i in range(B)]
    for shared in (False, True):
        outs = ctx.prove_batch(c, seeds, shared, S, phi, a, b, chs)
        for i in range(B):
            sd = seeds[0] if shared else seeds[i]
            rc, ref = orc.prove(co, sd, S[i], phi[i], a[i], b[i], chs[i], ntt=True, nthreads=8)
            assert rc == 0
            for k in ("t", "g", "u_1", "projection_int", "b_prime_prime", "h", "u_2", "z"):
                assert np.array_equal(outs[i][k], ref[k]), (shared, i, k)


def test_large_shape_sampled_rows(ctx, orc):
    """BASELINE config 4 shape (N = 2^16, R = 2^8, kappa = 2^22): sampled commitment rows, z and the exact witness norm
    against the oracle (the full T is 275 GB and never materialised, SURVEY 8d cfg 4).  Exercises R > 64 (several
    consumer passes), 64-bit counters and 4 GB operands."""
    N, R = 1 << 16, 1 << 8
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    dS = ctx.malloc(R * N * 256)
    ctx.synth_zq_dev(synth.SEED, 1, 0, R * N * 64, dS)
    ctx.witness_load_dev(c, dS)
    rows = [0, c.KAPPA - 1]
    S = np.empty((R, N, 64), np.uint32)
    ctx.d2h(S, dS); ctx.sync()
    assert np.array_equal(S[3, 77], synth.prg_zq(synth.SEED, 1, 64, start=(3 * N + 77) * 64))      # device PRG == host PRG
    dT = ctx.malloc(R * 1 * 256)
    for row in rows:
        ctx.commit_inner_dev(SEED32, row, 1, dT)
        T = np.empty((R, 1, 64), np.uint32)
        ctx.d2h(T, dT); ctx.sync()
        ref = orc.commit_inner_rows(co, SEED32, S, row, 1, ntt=True, nthreads=8)
        assert np.array_equal(T, ref), row
    assert ctx.norm_sq_dev(dS, R * N * 64) == int((S.astype(np.uint64) ** 2).sum())
    ctx.free(dS); ctx.free(dT)


def test_comm_single_rank_and_sharded_entry_points(ctx, orc):
    """lab_comm_* with a one-rank NCCL communicator (the multi-rank row sharding itself is exercised by
    `bench.py --workload prove` under torchrun, which checks that all ranks hold the same transcript digest as a
    single-GPU run), and the row-range forms the sharding is built from."""
    uid = lb.Context.comm_unique_id()
    assert len(uid) == 128
    c2 = lb.Context(0)
    try:
        c2.comm_init(uid, 0, 1)
        N, R = 2, 2
        c = lb.RuntimeConstants.new(N, R)
        co, _ = orc.constants(N, R)
        S = orc.generate_witness(co, 3)
        phi, a, b = orc.generate_state(co, S, 3)
        ch = orc.sample_challenges(co, 3, 2)
        rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=4)
        assert rc == 0
        st = lb.State(phi, a, b)
        ver = lb.Verifier.new(st.b_prime_k, c, challenges=ch)
        tr = lb.Prover.new(S, ver, c, c2).proof_gen(st, lb.CRS.from_seed(c, SEED32, c2)).as_oracle_dict()
        for k in ("t", "g", "u_1", "h", "u_2", "z"):
            assert np.array_equal(tr[k], ref[k]), k
        c2.comm_destroy()
    finally:
        c2.close()
    # row shards of T written with the destination stride of the full T, as the multi-GPU path does
    Tfull = ctx.commit_inner(c, SEED32, S)
    parts = [ctx.commit_inner(c, SEED32, S, row0=r0, nrows=32) for r0 in range(0, c.KAPPA, 32)]
    assert np.array_equal(np.concatenate(parts, axis=1), Tfull)


# ---- device-side generation (f2 challenge polys, f4 witness / statement) ----
def test_challenge_polys_on_device_match_oracle(ctx, orc):
    """Verifier::fetch_challenge (verification.rs:460-489) with the 1000-sample operator-norm rejection
    (util.rs:227-246): same draws, same f64 accept/reject decisions as the oracle and the host generator."""
    got, tries = ctx.sample_challenge_polys(synth.SEED, 0, 6)
    for i in range(6):
        assert np.array_equal(got[i], orc.sample_challenge_poly(synth.SEED, i)), i
    assert np.array_equal(got[2], synth.sample_challenge_poly(synth.SEED, 2))
    assert (tries >= 1).all()
    # shape of the challenge space: 23 zeros, 31 coefficients +-1, 10 coefficients +-2
    for p in got:
        v = np.where(p > Q // 2, Q - p.astype(np.int64), p.astype(np.int64))
        assert sorted(np.bincount(v, minlength=3).tolist()) == [10, 23, 31]
    other, _ = ctx.sample_challenge_polys(12345, 7, 3)
    for k in range(3):
        assert np.array_equal(other[k], orc.sample_challenge_poly(12345, 7 + k))


@pytest.mark.parametrize("N,R", [(1, 1), (2, 2), (3, 5), (8, 8)])
def test_generate_witness_and_state_on_device(ctx, orc, N, R):
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S, norm, draws = ctx.generate_witness(c, 77 + N)
    ref = orc.generate_witness(co, 77 + N)
    assert np.array_equal(S, ref)
    assert norm == int((ref.astype(np.uint64) ** 2).sum()) <= c.BETA_BOUND ** 2
    assert draws % 2 == 0 and draws > 0
    phi, a, b = ctx.generate_state(c, 5, S)
    rphi, ra, rb = orc.generate_state(co, S, 5)
    assert np.array_equal(phi, rphi) and np.array_equal(a, ra) and np.array_equal(b, rb)


def test_crs_cache_is_transparent(ctx, orc):
    """lab_crs_cache_configure: the first proof fills the cache (write-through in K_MV), verify and a second proof read
    it back; every result stays bit-identical to the uncached run and to the oracle."""
    N, R = 2, 3
    c = lb.RuntimeConstants.new(N, R)
    co, _ = orc.constants(N, R)
    S = orc.generate_witness(co, 21)
    phi, a, b = orc.generate_state(co, S, 21)
    ch = orc.sample_challenges(co, 21, 3)
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    st = lb.State(phi, a, b)
    ver = lb.Verifier.new(st.b_prime_k, c, challenges=ch)
    c2 = lb.Context(0)
    try:
        c2.crs_cache_configure(1 << 30)
        prover = lb.Prover.new(S, ver, c, c2)
        crs = lb.CRS.from_seed(c, SEED32, c2)
        tr1 = prover.proof_gen(st, crs).as_oracle_dict()
        s1 = c2.crs_cache_stats()
        assert s1["misses"] == 3 and s1["hits"] == 0 and s1["bytes"] > 0          # A, u_1 and u_2 generated and stored
        ok = c2.verify(c, SEED32, phi, a, b, ch, tr1)
        s2 = c2.crs_cache_stats()
        assert ok[0] and s2["hits"] == 3                                          # Checks 15 (A z), 19, 20 from the cache
        # the second proof of a small shape is recorded as a CUDA graph, which keeps its own copy of u_1's CRS side: generated by
        # this replay (a miss), reused by the next one under the same seed (a hit: the generation node is disabled)
        tr2 = prover.proof_gen(st, crs).as_oracle_dict()
        s3 = c2.crs_cache_stats()
        assert (s3["hits"], s3["misses"]) == (3, 4) and c2.graph_stats()["graphs"] == 1
        tr2b = prover.proof_gen(st, crs).as_oracle_dict()
        s4 = c2.crs_cache_stats()
        assert (s4["hits"], s4["misses"]) == (4, 4)
        for k in ("t", "g", "u_1", "projection_int", "b_prime_prime", "h", "u_2", "z"):
            assert np.array_equal(tr1[k], ref[k]), k
            assert np.array_equal(tr2[k], ref[k]), k
            assert np.array_equal(tr2b[k], ref[k]), k
        # a different seed must not hit
        other = bytes(range(1, 33))
        rc, ref_o = orc.prove(co, other, S, phi, a, b, ch, ntt=True, nthreads=8)
        tr3 = prover.proof_gen(st, lb.CRS.from_seed(c, other, c2)).as_oracle_dict()
        assert np.array_equal(tr3["u_1"], ref_o["u_1"]) and np.array_equal(tr3["u_2"], ref_o["u_2"])
        assert c2.crs_cache_stats()["misses"] == 5
        tr3b = prover.proof_gen(st, crs).as_oracle_dict()                          # back to the first seed: generated again, not served stale
        assert np.array_equal(tr3b["u_1"], ref["u_1"]) and c2.crs_cache_stats()["misses"] == 6
        # a tampered transcript is still rejected at the same check when the verifier reads the cache
        bad = dict(tr1); bad["u_1"] = np.array(tr1["u_1"], copy=True); bad["u_1"][0, 0] ^= 1
        assert c2.verify(c, SEED32, phi, a, b, ch, bad)[:2] == (False, 19)
        # the inner commitment alone, 64-vector shape (IC = 16 kernel): fill, then read
        Nw, Rw = 3, 40
        cw = lb.RuntimeConstants.new(Nw, Rw, allow_degenerate=True)
        cow, _ = orc.constants(Nw, Rw)
        Sw = synth.uniform_witness(Nw, Rw, seed=4)
        refT = orc.commit_inner_rows(cow, SEED32, Sw, 2, 21, nthreads=8)
        h0 = c2.crs_cache_stats()["hits"]
        assert np.array_equal(c2.commit_inner(cw, SEED32, Sw, row0=2, nrows=21), refT)
        assert np.array_equal(c2.commit_inner(cw, SEED32, Sw, row0=2, nrows=21), refT)
        assert c2.crs_cache_stats()["hits"] == h0 + 1
        c2.crs_cache_configure(0)
        assert c2.crs_cache_stats()["bytes"] == 0
        tr4 = prover.proof_gen(st, crs).as_oracle_dict()
        assert np.array_equal(tr4["u_1"], ref["u_1"])
    finally:
        c2.close()


@pytest.mark.parametrize("N,R,row0,nrows", [(3, 70, 1, 21), (5, 130, 0, 70), (16400, 65, 2, 3)])
def test_commit_inner_more_than_64_vectors_uses_one_chacha_pass(ctx, orc, N, R, row0, nrows):
    """R > 64: A is generated once per row chunk into transient int8 limb planes (first 64 vectors on the CUDA cores) and
    the remaining vectors are contracted on the tensor cores (lab_umma.cuh); N = 16400 needs two K-segments (s32
    accumulators).  Bit-exact against the oracle, which multiplies row by row."""
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=N + R)
    got = ctx.commit_inner(c, SEED32, S, row0=row0, nrows=nrows)
    ref = orc.commit_inner_rows(co, SEED32, S, row0, nrows, ntt=True, nthreads=8)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("N,R,row0,nrows", [(1, 1, 0, 5), (3, 2, 1, 21), (5, 130, 0, 70), (33, 64, 3, 130), (16400, 3, 2, 3), (33, 5, 2, 450)])
def test_commit_inner_generate_then_contract(ctx, orc, N, R, row0, nrows, monkeypatch):
    """The large-shape cold path (k_gen_planes: every warp regenerates A into int8 limb planes; k_umma_commit: tensor-core
    contraction with all witness vectors), forced on small shapes the oracle can check: ragged rows and columns (zero
    padding of planes), R = 1 (N' = 16), several passes of 64 vectors, two K-segments; 450 rows of N = 33 are four chunks, so
    the two plane buffers are reused while the contraction of chunk k runs on its own stream beside the generation of
    chunk k + 1 (LAB_GC_OVERLAP=0: one buffer, one stream -- same bits)."""
    monkeypatch.setenv("LAB_GEN_CONTRACT_MIN_POLYS", "1")
    monkeypatch.setenv("LAB_GC_CHUNK_MB", "1")          # 1 MB of limb planes per chunk: several row chunks, T streamed out per chunk
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=3 * N + R)
    got = ctx.commit_inner(c, SEED32, S, row0=row0, nrows=nrows)
    ref = orc.commit_inner_rows(co, SEED32, S, row0, nrows, ntt=True, nthreads=8)
    assert np.array_equal(got, ref)
    monkeypatch.setenv("LAB_GC_OVERLAP", "0")
    assert np.array_equal(ctx.commit_inner(c, SEED32, S, row0=row0, nrows=nrows), ref)
    monkeypatch.setenv("LAB_GEN_CONTRACT_MIN_POLYS", "0")          # 0 = never: the warp-specialised K_A path
    assert np.array_equal(ctx.commit_inner(c, SEED32, S, row0=row0, nrows=nrows), ref)


@pytest.mark.parametrize("low", [(5 << 32) + (1 << 32) - 64 * 5 - 20, (1 << 64) - 64 * 3 - 10])
def test_generate_then_contract_counter_boundaries(ctx, orc, low, monkeypatch):
    """k_gen_planes at the 2^32 boundary of seed + counter (hoisted ChaCha20 state changes inside polynomial 5) and at the
    2^64 carry into the upper seed limbs (inside polynomial 3): same seeds as the K_A boundary tests."""
    monkeypatch.setenv("LAB_GEN_CONTRACT_MIN_POLYS", "1")
    N, R = 4, 2
    seed = bytes(range(7, 31)) + low.to_bytes(8, "big")
    c = lb.RuntimeConstants.new(N, R)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=9)
    assert np.array_equal(ctx.commit_inner(c, seed, S, 0, 8), orc.commit_inner_rows(co, seed, S, 0, 8))


# ---- round 2: 2-bit packed JL matrices (lab_jl.cuh), sharded stage forms, BASELINE shapes, larger / non-square proofs ----
@pytest.mark.parametrize("N,R", [(1, 1), (2, 3), (7, 3), (8, 2), (9, 2), (33, 5), (70, 2), (33, 9), (70, 4), (129, 3), (256, 2)])
def test_packed_jl_matches_int8_and_oracle(ctx, orc, N, R):
    """lab_jl_project2 / lab_aggregate_phi2 on the packed matrices == the int8 entry points == the oracle (proofgen.rs:429-457,
    :244-253).  Up to 2^14 coefficients the thread-per-row kernel runs, above it the table kernel: N = 33, 70, 129 leave its
    last 512-coefficient unit ragged, (256, 2) fills its units exactly."""
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=31 * N + R)
    pi = synth.sample_pi(N, R, seed=9)
    pi2 = lb.api.pack_pi(pi)
    assert pi2.shape == (R, 256, N * 4)
    assert np.array_equal(lb.api.unpack_pi(pi2), pi)
    ref = orc.jl_project(co, S, pi)
    p8, acc8 = ctx.jl_project(c, S, pi)
    p2, acc2 = ctx.jl_project2(c, S, pi2)
    assert np.array_equal(p8, ref) and np.array_equal(p2, ref)
    assert acc8 == acc2 == orc.valid_projection(co, ref)
    h = R // 2
    parts = ctx.jl_project2_part(c, S, pi2[:h], 0, h) + ctx.jl_project2_part(c, S, pi2[h:], h, R - h)
    assert np.array_equal(parts, ref)
    # phi'' (S5): psi phi + sigma_inv(Pi^T omega), reference value from the definition in numpy
    phi = rand_polys(R * N, 12).reshape(R, N, D)
    omega = synth.prg_zq(5, 7, 256)
    psi = 4097
    v = (np.einsum("j,ijc->ic", omega.astype(np.int64), pi.astype(np.int64)) % Q).reshape(R, N, D)
    conj = np.empty_like(v)
    conj[..., 0] = v[..., 0]
    conj[..., 1:] = (Q - v[..., :0:-1]) % Q
    want = ((phi.astype(np.int64) * psi + conj) % Q).astype(np.uint32)
    assert np.array_equal(ctx.aggregate_phi(c, phi, pi, psi, omega), want)
    assert np.array_equal(ctx.aggregate_phi2(c, phi, pi2, psi, omega), want)


def test_packed_jl_extreme_values(ctx, orc):
    """All-plus / all-minus rows against a witness of maximal coefficients: the largest table entries (8 x 8190) and the
    largest per-lane sums; a non-canonical witness is reduced before it is projected (Zq values are canonical in the
    reference, algebraic.rs:30-38)."""
    N, R = 16, 2
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = np.full((R, N, D), Q - 1, np.uint32)
    pi = np.zeros((R, 256, N * D), np.int8)
    pi[:, 0::3, :] = 1
    pi[:, 1::3, :] = -1
    pi[1, 5, 100:] = 0
    ref = orc.jl_project(co, S, pi)
    assert np.array_equal(ctx.jl_project2(c, S, lb.api.pack_pi(pi))[0], ref)
    assert abs(int(ref[0])) == R * N * D * (Q - 1)
    Snc = S + np.uint32(3 * Q)                                # same residues, non-canonical representatives
    assert np.array_equal(ctx.jl_project(c, Snc, pi)[0], ref)
    with pytest.raises(lb.LabError):
        lb.api.pack_pi(np.full((1, 256, 64), 2, np.int8))


def test_full_proof_with_packed_pi(ctx, orc):
    """lab_prove / lab_verify / the bincode writer take the JL attempts 2-bit packed (lab_challenges.pi2): same transcript."""
    N, R = 2, 3
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=515, n_attempts=3)
    c = lb.RuntimeConstants.new(N, R)
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    ch2 = dict(ch); ch2["pi2"] = lb.api.pack_pi(ch["pi"]); ch2["pi"] = None
    st = lb.State(phi, a, b)
    tr = lb.Prover.new(S, lb.Verifier.new(st.b_prime_k, c, challenges=ch2), c, ctx).proof_gen(st, lb.CRS.from_seed(c, SEED32, ctx))
    got = tr.as_oracle_dict()
    for k in ("t", "g", "u_1", "projection_int", "projection", "b_prime_prime", "phi_final", "h", "u_2", "z"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(tr.pi_accepted, ch["pi"][ref["jl_attempt"]])
    assert ctx.verify(c, SEED32, phi, a, b, ch2, got)[:2] == (True, 0)
    assert lb.api.transcript_bincode(c, got, ch2) == lb.api.transcript_bincode(c, got, ch)


def test_stage_shards_add_up(ctx, orc):
    """lab_gram_part / lab_amortize_z_part: the (i, .) tiles of g and the per-shard partial sums of z."""
    N, R = 5, 7
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=55)
    ch = rand_polys(R, 13)
    G = orc.gram(co, S)
    assert np.array_equal(np.concatenate([ctx.gram_part(c, S, 0, 3), ctx.gram_part(c, S, 3, 4)]), G)
    z = sum(ctx.amortize_z_part(c, S, ch, i0, ni).astype(np.uint64) for i0, ni in ((0, 2), (2, 0), (2, 5))) % Q
    assert np.array_equal(z.astype(np.uint32), orc.amortize_z(co, S, ch))


def test_resident_witness_host_forms(ctx, orc, monkeypatch):
    """lab_witness_load + lab_commit_inner_resident (host destination, rows streamed out per chunk) + the sharded device
    calls without a communicator (they then compute everything locally)."""
    monkeypatch.setenv("LAB_GEN_CONTRACT_MIN_POLYS", "1")
    monkeypatch.setenv("LAB_GC_CHUNK_MB", "1")
    N, R = 33, 6
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    S = synth.uniform_witness(N, R, seed=8)
    c2 = lb.Context(0)
    try:
        c2.witness_load(c, S)
        T = np.empty((R, 200, D), np.uint32)
        import ctypes as C
        c2._ck(c2.L.lab_commit_inner_resident(c2._h, lb.api._p(lb.api._seed_buf(SEED32)), C.c_uint64(5), C.c_uint64(200), lb.api._p(T)))
        assert np.array_equal(T, orc.commit_inner_rows(co, SEED32, S, 5, 200, ntt=True, nthreads=8))
        pi = synth.sample_pi(N, R, seed=2)
        pi2 = lb.api.pack_pi(pi)
        dpi, dp, dG, dz, dch = c2.malloc(pi2.nbytes), c2.malloc(256 * 8), c2.malloc(R * R * 256), c2.malloc(N * 256), c2.malloc(R * 256)
        chal = rand_polys(R, 14)
        c2.h2d(dpi, pi2); c2.h2d(dch, chal)
        assert c2.comm_shard(R) == (0, R)
        c2.jl_project_sharded_dev(dpi, dp)
        c2.gram_sharded_dev(dG)
        c2.amortize_z_sharded_dev(dch, dz)
        p, G, z = np.empty(256, np.int64), np.empty((R, R, D), np.uint32), np.empty((N, D), np.uint32)
        c2.d2h(p, dp); c2.d2h(G, dG); c2.d2h(z, dz); c2.sync()
        assert np.array_equal(p, orc.jl_project(co, S, pi))
        assert np.array_equal(G, orc.gram(co, S))
        assert np.array_equal(z, orc.amortize_z(co, S, chal))
        for d in (dpi, dp, dG, dz, dch):
            c2.free(d)
    finally:
        c2.close()


@pytest.mark.parametrize("N,R", [(8, 8), (2, 4), (4, 8)])
def test_full_proof_larger_and_non_square_shapes(ctx, orc, N, R):
    """labrador_perf shapes beyond (4,4) (benches/labrador_perf.rs:22-28 doubles n and r alternately from (1,2)): (8,8) has
    T_1 = 5, (2,4) and (4,8) are the non-square steps.  Transcript equal field by field, oracle verifier accepts."""
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=2000 + 10 * N + R, n_attempts=3)
    c = lb.RuntimeConstants.new(N, R)
    nth = orc.num_threads()
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=nth)
    assert rc == 0
    st = lb.State(phi, a, b)
    tr = lb.Prover.new(S, lb.Verifier.new(st.b_prime_k, c, challenges=ch), c, ctx).proof_gen(st, lb.CRS.from_seed(c, SEED32, ctx))
    got = tr.as_oracle_dict()
    for k in ("t", "g", "u_1", "projection_int", "projection", "b_prime_prime", "phi_final", "h", "u_2", "z"):
        assert np.array_equal(got[k], ref[k]), k
    assert got["jl_attempt"] == ref["jl_attempt"]
    ok, failed, norm_sum = orc.verify(co, SEED32, phi, a, b, ch, got, ntt=True, nthreads=nth)
    assert ok and failed == 0 and tr.norm_sum == norm_sum
    assert ctx.verify(c, SEED32, phi, a, b, ch, got) == (True, 0, norm_sum)


@pytest.mark.parametrize("N,R", [(1, 2), (2, 2), (2, 4)])
def test_full_proof_against_schoolbook_oracle(ctx, orc, N, R):
    """The same comparison with the oracle multiplying by the reference's classic path (schoolbook product + sign-folding
    reduction, algebraic.rs:352-376,402: NTT_ENABLED = false) -- the checker then shares no transform with the device."""
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=3000 + 10 * N + R, n_attempts=3)
    c = lb.RuntimeConstants.new(N, R)
    nth = orc.num_threads()
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=False, nthreads=nth)
    assert rc == 0
    st = lb.State(phi, a, b)
    got = lb.Prover.new(S, lb.Verifier.new(st.b_prime_k, c, challenges=ch), c, ctx).proof_gen(st, lb.CRS.from_seed(c, SEED32, ctx)).as_oracle_dict()
    for k in ("t", "g", "u_1", "projection_int", "projection", "b_prime_prime", "phi_final", "h", "u_2", "z"):
        assert np.array_equal(got[k], ref[k]), k
    ok, failed, _ = orc.verify(co, SEED32, phi, a, b, ch, got, ntt=False, nthreads=nth)
    assert ok and failed == 0


def test_rejected_proof_then_other_shape_on_same_ctx(ctx, orc):
    """A proof that ends with LAB_ERR_JL_REJECTED leaves its forked strand (inner commitment, g, u_1 on the second stream)
    behind; the next call on the same ctx -- another shape, so other arena offsets -- must not run into it."""
    N, R = 16, 16
    co, _ = orc.constants(N, R)
    c = lb.RuntimeConstants.new(N, R)
    S_big = synth.uniform_witness(N, R, seed=3)              # not short: every projection is rejected
    phi, a = synth.generate_statement_inputs(N, R, 5)
    st = lb.State(phi, a, np.zeros(D, np.uint32))
    ch = {"pi": np.stack([synth.sample_pi(N, R, 7, t) for t in range(6)]), "psi": 1, "omega": synth.prg_zq(7, 7, 256), "alpha": synth.prg_zq(7, 8, D),
          "beta": synth.prg_zq(7, 9, D), "c": rand_polys(R, 15)}
    c2 = lb.Context(0)
    try:
        with pytest.raises(lb.LabError) as e:
            lb.Prover.new(S_big, lb.Verifier.new(st.b_prime_k, c, challenges=ch), c, c2).proof_gen(st, lb.CRS.from_seed(c, SEED32, c2))
        assert e.value.status == 1
        # immediately afterwards: a small proof of a different shape on the same context, checked against the oracle
        n2, r2 = 2, 3
        co2, S2, phi2, a2, b2, ch2 = full_case(orc, n2, r2, seed=808)
        cc2 = lb.RuntimeConstants.new(n2, r2)
        rc, ref = orc.prove(co2, SEED32, S2, phi2, a2, b2, ch2, ntt=True, nthreads=8)
        assert rc == 0
        st2 = lb.State(phi2, a2, b2)
        got = lb.Prover.new(S2, lb.Verifier.new(st2.b_prime_k, cc2, challenges=ch2), cc2, c2).proof_gen(st2, lb.CRS.from_seed(cc2, SEED32, c2)).as_oracle_dict()
        for k in ("t", "g", "u_1", "h", "u_2", "z"):
            assert np.array_equal(got[k], ref[k]), k
    finally:
        c2.close()


def test_baseline_cfg3_shape_against_oracle(ctx, orc):
    """BASELINE config 3 at its full size, (N, R) = (4096, 64), kappa = 262144, through the default code paths (no environment
    overrides): the whole JL projection (4.3e9 entries) against the oracle's literal loop, the amortised opening z, 64
    sampled garbage polynomials g_ij, and rows of T around the first 4 GB chunk boundary of the generate-then-contract
    path (8191 / 8192 / 8193), rows 0, 1 and the last row kappa - 1.  One of the rows is also checked with the oracle's
    schoolbook multiplication."""
    N, R = 4096, 64
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    co, _ = orc.constants(N, R)
    ND, kappa = N * D, c.KAPPA
    nth = orc.num_threads()
    c2 = lb.Context(0)
    try:
        dS = c2.malloc(R * ND * 4)
        c2.synth_zq_dev(synth.SEED, 1, 0, R * ND, dS)
        c2.witness_load_dev(c, dS)
        S = np.empty((R, N, D), np.uint32)
        c2.d2h(S, dS); c2.sync()
        assert np.array_equal(S[5, 9], synth.prg_zq(synth.SEED, 1, D, start=(5 * N + 9) * D))
        # ---- G4: full projection, packed on the device from the int8 stream and generated packed directly ----
        dpi8, dpi2, dpi2b, dp = c2.malloc(R * 256 * ND), c2.malloc(R * 256 * ND // 4), c2.malloc(R * 256 * ND // 4), c2.malloc(256 * 8)
        c2.synth_pi_dev(synth.SEED, 0, 0, R * 256 * ND, dpi8)
        c2.synth_pi2_dev(synth.SEED, 0, 0, R * 256 * ND, dpi2b)
        c2.pi_pack_dev(dpi8, R * 256 * ND, dpi2)
        pi = np.empty((R, 256, ND), np.int8)
        c2.d2h(pi, dpi8)
        a2, b2 = np.empty(R * 256 * ND // 16, np.uint32), np.empty(R * 256 * ND // 16, np.uint32)
        c2.d2h(a2, dpi2); c2.d2h(b2, dpi2b); c2.sync()
        assert np.array_equal(a2, b2)
        del a2, b2
        p = np.empty(256, np.int64)
        c2.jl_project2_dev(dpi2, 0, R, dp)
        c2.d2h(p, dp); c2.sync()
        ref_p = orc.jl_project(co, S, pi)
        assert np.array_equal(p, ref_p)
        h = 24                                               # two unequal shards of witness vectors: partial sums add up
        c2.jl_project2_dev(dpi2, 0, h, dp); p0 = np.empty(256, np.int64); c2.d2h(p0, dp); c2.sync()
        c2.jl_project2_dev(dpi2 + h * 256 * ND // 4, h, R - h, dp); p1 = np.empty(256, np.int64); c2.d2h(p1, dp); c2.sync()
        assert np.array_equal(p0 + p1, ref_p)
        c2.jl_project_dev(dpi8, 0, R, dp); c2.d2h(p0, dp); c2.sync()          # int8 entry point (packs first)
        assert np.array_equal(p0, ref_p)
        del pi
        for d in (dpi8, dpi2, dpi2b):
            c2.free(d)
        # ---- G9: z ----
        chal = rand_polys(R, 16)
        dch, dz = c2.malloc(R * 256), c2.malloc(N * 256)
        c2.h2d(dch, chal)
        c2.amortize_z_dev(dch, 0, R, dz)
        z = np.empty((N, D), np.uint32)
        c2.d2h(z, dz); c2.sync()
        assert np.array_equal(z, orc.amortize_z(co, S, chal))
        # ---- G2: sampled g_ij ----
        dG = c2.malloc(R * R * 256)
        c2.gram_dev(0, R, dG)
        G = np.empty((R, R, D), np.uint32)
        c2.d2h(G, dG); c2.sync()
        assert np.array_equal(G, G.transpose(1, 0, 2))
        rng = np.random.default_rng(3)
        for i, j in [(0, 0), (63, 63), (0, 63)] + [tuple(rng.integers(0, R, 2)) for _ in range(61)]:
            assert np.array_equal(G[i, j], orc.inner_product(S[i], S[j])), (i, j)
        # ---- G1: rows of T around the chunk boundary and at both ends (default generate-then-contract path) ----
        n0 = 8448
        dT = c2.malloc(R * n0 * 256)
        c2.commit_inner_dev(SEED32, 0, n0, dT)
        T = np.empty((R, n0, D), np.uint32)
        c2.d2h(T, dT); c2.sync()
        for row in (0, 1, 8191, 8192, 8193, n0 - 1):
            ref = orc.commit_inner_rows(co, SEED32, S, row, 1, ntt=True, nthreads=nth)
            assert np.array_equal(T[:, row:row + 1], ref), row
        assert np.array_equal(T[:, 8192:8193], orc.commit_inner_rows(co, SEED32, S, 8192, 1, ntt=False, nthreads=nth))
        n1 = 1024
        c2.commit_inner_dev(SEED32, kappa - n1, n1, dT)
        T1 = np.empty((R, n1, D), np.uint32)
        c2.d2h(T1, dT); c2.sync()
        for row in (kappa - n1, kappa - 1):
            ref = orc.commit_inner_rows(co, SEED32, S, row, 1, ntt=True, nthreads=nth)
            assert np.array_equal(T1[:, row - (kappa - n1):row - (kappa - n1) + 1], ref), row
        for d in (dS, dp, dch, dz, dG, dT):
            c2.free(d)
    finally:
        c2.close()


def test_graph_path_matches_plain_path(orc, monkeypatch):
    """Small shapes replay the whole proof as one CUDA graph from the second proof of a shape on (lab_api.cu, prove_graph): new
    CRS seed, statement, challenges and psi per replay, int8 and packed matrices (separate graphs), a rejected first JL
    attempt (the graph is replayed with the next matrices), and the same proofs again with LAB_NO_GRAPH on a fresh context."""
    N, R = 2, 2
    c = lb.RuntimeConstants.new(N, R)
    cases = []
    for k in range(5):
        co, S, phi, a, b, ch = full_case(orc, N, R, seed=9000 + k, n_attempts=3)
        if k == 3:                                   # attempt 0 = all-ones rows: p_j = sum of the coefficients for every j -> rejected
            ch["pi"] = ch["pi"].copy(); ch["pi"][0] = 1
            assert not orc.valid_projection(co, orc.jl_project(co, S, ch["pi"][0]))
        seed32 = bytes([k + 1]) * 32
        rc, ref = orc.prove(co, seed32, S, phi, a, b, ch, ntt=True, nthreads=8)
        assert rc == 0
        cases.append((seed32, S, phi, a, b, ch, ref))
    assert cases[3][6]["jl_attempt"] == 1

    def run_all(cx, packed):
        launches = []
        for seed32, S, phi, a, b, ch, ref in cases:
            chx = dict(ch)
            if packed:
                chx["pi2"] = lb.api.pack_pi(ch["pi"]); chx["pi"] = None
            st = lb.State(phi, a, b)
            l0 = cx.kernel_launches
            tr = lb.Prover.new(S, lb.Verifier.new(st.b_prime_k, c, challenges=chx), c, cx).proof_gen(st, lb.CRS.from_seed(c, seed32, cx))
            launches.append(cx.kernel_launches - l0)
            got = tr.as_oracle_dict()
            for k in ("t", "g", "u_1", "projection_int", "projection", "b_prime_prime", "phi_final", "h", "u_2", "z"):
                assert np.array_equal(got[k], ref[k]), (packed, k)
            assert got["jl_attempt"] == ref["jl_attempt"]
            assert tr.norm_sum == orc.verify(orc.constants(N, R)[0], seed32, phi, a, b, ch, got, ntt=True, nthreads=8)[2]
        return launches
    c2 = lb.Context(0)
    try:
        run_all(c2, False)
        assert c2.graph_stats() == {"graphs": 1, "replays": 5, "failed": False}      # 4 proofs after the first + 1 replay for the rejected attempt
        run_all(c2, True)
        assert c2.graph_stats() == {"graphs": 2, "replays": 10, "failed": False}
    finally:
        c2.close()
    monkeypatch.setenv("LAB_NO_GRAPH", "1")
    c3 = lb.Context(0)
    try:
        run_all(c3, False)
        assert c3.graph_stats()["graphs"] == 0
    finally:
        c3.close()


@pytest.mark.parametrize("N,R", [(1, 2), (2, 2), (2, 3)])
def test_fiat_shamir_proof_matches_oracle(ctx, orc, N, R):
    """lab_prove_fs: the challenges come from the SHA-256 chain over the transcript prefix (include/labrador_b200.h).  The
    independent restatement of the chain (labrador_b200/fs.py: hashlib + the host samplers) derives the same challenges from
    the GPU transcript; the CPU oracle, given them by injection, reproduces every field; the oracle's verifier and lab_verify_fs
    accept; a tampered transcript is rejected (the derived challenges no longer fit, or a check fails)."""
    co, rc = orc.constants(N, R)
    c = lb.RuntimeConstants.new(N, R)
    S = orc.generate_witness(co, 600 + N + R)
    phi, a, b = orc.generate_state(co, S, 600 + N + R)
    tr, chg = ctx.prove_fs(c, SEED32, S, phi, a, b)
    ch = lb.fs.derive_challenges(N, R, SEED32, phi, a, b, tr)
    att = tr["jl_attempt"]
    assert chg["psi"] == ch["psi"]
    for k in ("omega", "alpha", "beta", "c"):
        assert np.array_equal(chg[k], ch[k]), k
    assert np.array_equal(lb.api.unpack_pi(chg["pi2"][0]), ch["pi"][att])
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0 and ref["jl_attempt"] == att
    for k in ("t", "g", "u_1", "projection_int", "projection", "b_prime_prime", "phi_final", "h", "u_2", "z"):
        assert np.array_equal(tr[k], ref[k]), k
    ok, failed, norm_sum = orc.verify(co, SEED32, phi, a, b, ch, tr, ntt=True, nthreads=8)
    assert ok and norm_sum == tr["norm_sum"]
    assert ctx.verify_fs(c, SEED32, phi, a, b, tr) == (True, 0, norm_sum)
    for field, idx in (("u_1", (0, 0)), ("z", (0, 3)), ("u_2", (1, 1)), ("b_prime_prime", (5,)), ("t", (0, 2, 2))):
        bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in tr.items()}
        bad[field][idx] = (int(bad[field][idx]) + 1) % Q
        assert ctx.verify_fs(c, SEED32, phi, a, b, bad)[0] is False, field
    bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in tr.items()}
    bad["projection_int"][0] += 10 ** 6                              # breaks the norm bound the verifier now checks itself
    assert ctx.verify_fs(c, SEED32, phi, a, b, bad)[:2] == (False, 7)
    # the compact wire format carries exactly what the non-interactive verifier needs besides the statement
    got, _ = lb.api.transcript_unpack(c, lb.api.transcript_pack(c, tr, chg))
    got["projection_int"] = tr["projection_int"]
    assert ctx.verify_fs(c, SEED32, phi, a, b, got)[0] is True


def test_rq_add_sub(ctx):
    """&Rq + &Rq and &Rq - &Rq (algebraic.rs:441-515): coefficientwise mod q, also on non-canonical input."""
    a = np.concatenate([rand_polys(200, 21), edge_polys()])
    b = np.concatenate([rand_polys(200, 22), edge_polys()[::-1]])
    assert np.array_equal(ctx.rq_add_batch(a, b), ((a.astype(np.uint64) + b) % Q).astype(np.uint32))
    assert np.array_equal(ctx.rq_add_batch(a, b, sub=True), ((a.astype(np.int64) - b) % Q).astype(np.uint32))
    assert np.array_equal(ctx.rq_add_batch(a + np.uint32(5 * Q), b), ((a.astype(np.uint64) + b) % Q).astype(np.uint32))


def test_cpp_header_runs_a_proof(ctx, orc, tmp_path):
    """tests/cpp/test_labrador_hpp.cpp: Prover::proof_gen / verify / to_bincode / size_in_bytes through cpp/labrador.hpp (compiled
    here with g++) against the oracle's transcript of the same inputs, int8 and packed challenges, a tampered u_2 -> check 20."""
    from test_host import build_cpp_test
    exe = build_cpp_test(tmp_path)
    N, R = 2, 2
    co, S, phi, a, b, ch = full_case(orc, N, R, seed=7100, n_attempts=2)
    c = lb.RuntimeConstants.new(N, R)
    rc, ref = orc.prove(co, SEED32, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    blob = lb.api.transcript_bincode(c, ref, ch)
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        f.write(np.array([N, R, 2, ref["jl_attempt"]], np.uint64).tobytes())
        f.write(SEED32)
        for arr in (S, phi, a, b):
            f.write(np.ascontiguousarray(arr, np.uint32).tobytes())
        f.write(np.ascontiguousarray(ch["pi"], np.int8).tobytes())
        f.write(np.array([ch["psi"]], np.uint32).tobytes())
        for k in ("omega", "alpha", "beta", "c"):
            f.write(np.ascontiguousarray(ch[k], np.uint32).tobytes())
        for k in ("u_1", "u_2", "z", "t", "g", "h"):
            f.write(np.ascontiguousarray(ref[k], np.uint32).tobytes())
        f.write(np.ascontiguousarray(ref["projection_int"], np.int64).tobytes())
        f.write(np.ascontiguousarray(ref["b_prime_prime"], np.uint32).tobytes())
        f.write(np.array([len(blob)], np.uint64).tobytes())
    import subprocess
    out = subprocess.run([exe, str(case)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_aggregate_phi_large_shape_uses_thread_per_word_kernel(ctx):
    """lab_aggregate_phi2 at (N, R) = (512, 32): 65536 packed words, the shape from which Pi^T omega runs one thread per word
    (k_piT_omega2) instead of one warp per word; reference = the definition in numpy, vector by vector."""
    N, R = 512, 32
    c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
    dpi = ctx.malloc(R * 256 * N * D // 4)
    ctx.synth_pi2_dev(synth.SEED, 3, 0, R * 256 * N * D, dpi)
    pi2 = np.empty((R, 256, N * 4), np.uint32)
    ctx.d2h(pi2, dpi); ctx.sync()
    ctx.free(dpi)
    phi = rand_polys(R * N, 18).reshape(R, N, D)
    omega = synth.prg_zq(9, 7, 256)
    psi = 8000
    got = ctx.aggregate_phi2(c, phi, pi2, psi, omega)
    for i in (0, 7, R - 1):
        pi = lb.api.unpack_pi(pi2[i]).astype(np.int64)                      # [256][N*64]
        v = (omega.astype(np.int64) @ pi % Q).reshape(N, D)
        conj = np.empty_like(v)
        conj[:, 0] = v[:, 0]
        conj[:, 1:] = (Q - v[:, :0:-1]) % Q
        want = ((phi[i].astype(np.int64) * psi + conj) % Q).astype(np.uint32)
        assert np.array_equal(got[i], want), i
