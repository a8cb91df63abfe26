"""The reference's own property tests (tests/proptest.rs:14-81, 50 cases each) mirrored on the oracle, plus the
self-consistency checks the reference relies on (prove -> verify accepts, main.rs:106-107)."""
import numpy as np
import pytest

D, Q = 64, 8191
rng = np.random.default_rng(20261018)


def rand_poly(n=1):
    return rng.integers(0, Q, size=(n, D), dtype=np.uint32)


def test_ntt_preserves_result(orc):          # proptest.rs:14-24
    for _ in range(50):
        a, b = rand_poly()[0], rand_poly()[0]
        assert np.array_equal(orc.rq_mul(a, b, ntt=True), orc.rq_mul(a, b, ntt=False))
    for a in (np.zeros(D, np.uint32), np.full(D, Q - 1, np.uint32)):
        assert np.array_equal(orc.rq_mul(a, a, ntt=True), orc.rq_mul(a, a, ntt=False))


def test_ntt_roundtrip_and_slot_semantics(orc):
    import pyref  # noqa: F401
    exps = orc.ntt_slot_exponents()
    assert sorted(exps) == list(range(1, 128, 4))          # all e = 1 mod 4
    a = rand_poly()[0]
    f = orc.ntt_fwd(a)
    assert np.array_equal(orc.ntt_inv(f), a)
    # slot j = f(zeta^e_j) in F_Q[i], zeta = 2620 + 936 i
    def cmul(x, y):
        return ((x[0] * y[0] - x[1] * y[1]) % Q, (x[0] * y[1] + x[1] * y[0]) % Q)
    def cpow(x, e):
        r = (1, 0)
        for _ in range(e):
            r = cmul(r, x)
        return r
    for j in (0, 7, 31):
        pt = cpow((2620, 936), exps[j])
        acc, pw = (0, 0), (1, 0)
        for d in range(D):
            acc = ((acc[0] + int(a[d]) * pw[0]) % Q, (acc[1] + int(a[d]) * pw[1]) % Q)
            pw = cmul(pw, pt)
        assert (int(f[2 * j]), int(f[2 * j + 1])) == acc


@pytest.mark.parametrize("ntt", [False, True])
def test_linearity_of_inner_product(orc, ntt):   # proptest.rs:37-64
    for _ in range(50):
        a, b = rand_poly(16), rand_poly(16)
        c = int(rng.integers(0, Q))
        ab = orc.inner_product(a, b)
        cb = ((b.astype(np.uint64) * c) % Q).astype(np.uint32)
        assert np.array_equal(orc.inner_product(a, cb), ((ab.astype(np.uint64) * c) % Q).astype(np.uint32))


def test_sigma_inv_invariant(orc):               # proptest.rs:68-81
    for _ in range(50):
        a, b = rand_poly(16), rand_poly(16)
        inv_a = np.stack([orc.sigma_inv(p) for p in a])
        prod = orc.inner_product(inv_a, b)
        assert int(prod[0]) == int((a.astype(np.uint64) * b).sum() % Q)


def test_decompose_closed_form_equals_literal_loop(orc):
    """the closed form equals the literal Zq-operator loop (util.rs:389-442) for every coefficient
    value and bases 2..173."""
    import pyref
    allc = np.arange(Q, dtype=np.uint32)
    pad = (-Q) % D
    polys = np.concatenate([allc, np.zeros(pad, np.uint32)]).reshape(-1, D)
    for base in list(range(2, 40)) + [64, 90, 91, 128, 173]:
        exp = 1
        while base ** exp < Q:
            exp += 1
        for p in polys[:: max(1, len(polys) // 16)]:
            lit = orc.decompose(p, base, exp, literal=True)
            assert np.array_equal(lit, orc.decompose(p, base, exp))
            assert np.array_equal(lit, pyref.decompose_literal(p, base, exp))
    # digits beyond exp are dropped, missing digits are zero
    p = np.zeros(D, np.uint32); p[0] = 7; p[1] = 8190
    assert orc.decompose(p, 10, 1, literal=True)[0, 0] == 3      # 7 base 10 -> [3] (no carry)
    assert orc.decompose(p, 10, 6, literal=True)[5, 1] == 0


@pytest.mark.parametrize("N,R", [(1, 1), (2, 2)])
def test_prove_verify_roundtrip_and_tamper(orc, N, R):
    c, rc = orc.constants(N, R)
    assert rc == 0
    seed = bytes(range(32))
    S = orc.generate_witness(c, 5)
    assert orc.norm_sq(S) <= c.BETA_BOUND ** 2                 # generate_witness post-condition (proofgen.rs:480)
    phi, a, b = orc.generate_state(c, S, 5)
    ch = orc.sample_challenges(c, 5, 2)
    rc, tr = orc.prove(c, seed, S, phi, a, b, ch, ntt=True, nthreads=8)
    assert rc == 0
    ok, failed, norm = orc.verify(c, seed, phi, a, b, ch, tr, ntt=True, nthreads=8)
    assert ok and failed == 0 and norm <= c.BETA_PRIME
    # every check the verifier performs can fire
    for field, check in (("g", 8), ("z", 15), ("u_1", 19), ("u_2", 20)):
        bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in tr.items()}
        idx = (0, 1, 0) if field == "g" and R > 1 else tuple([0] * bad[field].ndim)
        bad[field][idx] ^= 1
        ok, failed, _ = orc.verify(c, seed, phi, a, b, ch, bad, ntt=True, nthreads=8)
        assert not ok
        if not (field == "g" and R == 1):
            assert failed == check, (field, failed)
    # a wrong statement b breaks the prover-side assert (verification.rs:550)
    b2 = b.copy(); b2[0] = (int(b2[0]) + 1) % Q
    rc, _ = orc.prove(c, seed, S, phi, a, b2, ch, ntt=True, nthreads=8)
    assert rc == 2


def test_independent_python_restatement_agrees(orc):
    import pyref
    N, R = 1, 1
    c, _ = orc.constants(N, R)
    seed = bytes(range(32))
    S = orc.generate_witness(c, 21)
    phi, a, b = orc.generate_state(c, S, 21)
    ch = orc.sample_challenges(c, 21, 2)
    rc, tr = orc.prove(c, seed, S, phi, a, b, ch)
    assert rc == 0
    tr2 = pyref.prove(N, R, dict(B_1=c.B_1, T_1=c.T_1, B_2=c.B_2, T_2=c.T_2, BETA_BOUND=c.BETA_BOUND), seed, S, phi, a, b, ch)
    for k in tr2:
        assert np.array_equal(tr[k], tr2[k]), k
