"""Seeded synthetic inputs and verifier randomness.

The reference is unseeded (thread_rng everywhere); parity needs injected inputs.  This is a
counter-based SplitMix64 PRG with seed 0x4C61425241444F52 ("LaBRADOR") and one stream id per tensor:
  1 witness coefficients, 2 witness reduction picks, 3 a_ij, 4 phi, 5 Pi (+ attempt << 8), 6 psi,
  7 omega, 8 alpha, 9 beta, 10 challenge polys (+ idx << 8), 11 operator-norm samples (+ idx << 8).
Distributions follow the reference; tests check this module against the oracle's own
generators, so the product never needs to import oracle/.
"""
import math

import numpy as np

D, Q, JL = 64, 8191, 256
SEED = 0x4C61425241444F52
_M64 = (1 << 64) - 1
_G = 0x9E3779B97F4A7C15
_S = 0xD1342543DE82EF95


def prg_u64(seed, stream, idx):
    """idx: int or uint64 array -> uint64 (array)."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed + stream * _S) & _M64)
        z = base + (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * np.uint64(_G)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _mulhi(x, n):
    """floor(x * n / 2^64) for uint64 array x and n < 2^32."""
    hi, lo = x >> np.uint64(32), x & np.uint64(0xFFFFFFFF)
    return (hi * np.uint64(n) + ((lo * np.uint64(n)) >> np.uint64(32))) >> np.uint64(32)


def prg_zq(seed, stream, n, start=0):
    idx = np.arange(start, start + n, dtype=np.uint64)
    return _mulhi(prg_u64(seed, stream, idx), Q).astype(np.uint32)


def prg_below(seed, stream, idx, n):
    """single draw in [0, n) (n may exceed 2^32)."""
    return (int(prg_u64(seed, stream, int(idx))) * int(n)) >> 64


def uniform_witness(N, R, seed=SEED):
    """W-uni: i.i.d. uniform coefficients, [R][N][64]."""
    return prg_zq(seed, 1, R * N * D).reshape(R, N, D)


def generate_witness(N, R, beta_bound, seed=SEED):
    """W-ref: generate_witness semantics (proofgen.rs:460-518): uniform, then floor-halve random polys
    until the sum of canonical squared norms is <= beta^2."""
    S = uniform_witness(N, R, seed).copy()
    norm = int((S.astype(np.uint64) ** 2).sum())
    bound = int(beta_bound) ** 2
    draw = 0
    while norm > bound:
        n = prg_below(seed, 2, draw, N)
        i = prg_below(seed, 2, draw + 1, R)
        draw += 2
        p = S[i, n]
        before = int((p.astype(np.uint64) ** 2).sum())
        p //= 2
        norm -= before - int((p.astype(np.uint64) ** 2).sum())
    return S


def generate_statement_inputs(N, R, seed=SEED):
    """Symmetric uniform a [R][R][64] and uniform phi [R][N][64] (structs.rs:289-318)."""
    a = np.zeros((R, R, D), dtype=np.uint32)
    for i in range(R):
        for j in range(i, R):
            v = prg_zq(seed, 3, D, start=(i * R + j) * D)
            a[i, j] = v
            a[j, i] = v
    phi = prg_zq(seed, 4, R * N * D).reshape(R, N, D)
    return phi, a


def sample_pi(N, R, seed=SEED, attempt=0):
    """JL matrices, entries {-1,0,1} with P = (1/4,1/2,1/4) (verification.rs:553-566): two PRG bits per
    entry, 00 -> -1, 11 -> +1, else 0; row-major fill order.  int8 [R][256][N*64]."""
    total = R * JL * N * D
    words = prg_u64(seed, 5 + (attempt << 8), np.arange((total + 31) // 32, dtype=np.uint64))
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    two = ((words[:, None] >> shifts) & np.uint64(3)).reshape(-1)[:total]
    out = np.zeros(total, dtype=np.int8)
    out[two == 0] = -1
    out[two == 3] = 1
    return out.reshape(R, JL, N * D)


def _negacyclic(a, b):
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))
    full = np.concatenate([full, np.zeros(2 * D - full.size, dtype=np.int64)])
    return (full[:D] - full[D:]) % Q


def sample_challenge_poly(seed, idx):
    """fetch_challenge (verification.rs:460-489): coefficients drawn without replacement from
    {0 x23, 1 x31, 2 x10}, nonzero ones negated with probability 1/2; resampled while the 1000-sample
    operator-norm estimate exceeds T = 15 (util.rs:83-104,227-246)."""
    draw = 0
    od = 0
    st, so = 10 + (idx << 8), 11 + (idx << 8)
    while True:
        dist = [0] * 23 + [1] * 31 + [2] * 10
        c = np.zeros(D, dtype=np.int64)
        for d in range(D):
            ri = prg_below(seed, st, draw, len(dist))
            coeff = dist.pop(ri)
            sgn = int(prg_u64(seed, st, draw + 1)) >> 63
            draw += 2
            c[d] = (Q - coeff) if (coeff > 0 and sgn) else coeff
        r = prg_zq(seed, so, 1000 * D, start=od).reshape(1000, D).astype(np.int64)
        od += 1000 * D
        sup = 0.0
        for s in range(1000):
            cr = _negacyclic(c, r[s])
            ratio = math.sqrt(float(int((cr * cr).sum()))) / math.sqrt(float(int((r[s] * r[s]).sum())))
            if ratio > sup:
                sup = ratio
        if not sup > 15.0:
            return c.astype(np.uint32)


def sample_challenges(N, R, seed=SEED, n_attempts=1):
    """Verifier randomness in the order proof_gen consumes it."""
    return {
        "pi": np.stack([sample_pi(N, R, seed, a) for a in range(n_attempts)]),
        "psi": int(prg_zq(seed, 6, 1)[0]),
        "omega": prg_zq(seed, 7, JL),
        "alpha": prg_zq(seed, 8, D),
        "beta": prg_zq(seed, 9, D),
        "c": np.stack([sample_challenge_poly(seed, i) for i in range(R)]),
    }
