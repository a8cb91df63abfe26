"""Independent numpy / big-int restatement used to cross-check the C oracle (TEST INFRASTRUCTURE).

Written separately from labrador_oracle.c on purpose: vectorised ChaCha20 over many keys, Python
big integers for the 256-bit counter and the 256-bit widening multiply of the rand-0.8.5 sampler,
numpy convolutions for ring products, and the proof_gen formulas of SURVEY appendix A.2 written
as array expressions.  Reference citations are the same as in labrador_oracle.c.
"""
import numpy as np

D, Q, JL = 64, 8191, 256
M32 = 0xFFFFFFFF


def _rotl(x, n):
    return ((x << np.uint32(n)) | (x >> np.uint32(32 - n))).astype(np.uint32)


def chacha20_blocks(keys, counter=0, stream=0):
    """keys: (n,8) uint32 -> (n,16) uint32 keystream block `counter` (64-bit ctr, 64-bit stream)."""
    keys = np.asarray(keys, dtype=np.uint32).reshape(-1, 8)
    n = keys.shape[0]
    s = np.zeros((16, n), dtype=np.uint32)
    s[0], s[1], s[2], s[3] = 0x61707865, 0x3320646E, 0x79622D32, 0x6B206574
    s[4:12] = keys.T
    s[12], s[13] = counter & M32, (counter >> 32) & M32
    s[14], s[15] = stream & M32, (stream >> 32) & M32
    x = s.copy()

    def qr(a, b, c, d):
        x[a] += x[b]; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] += x[d]; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] += x[b]; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] += x[d]; x[b] = _rotl(x[b] ^ x[c], 7)

    with np.errstate(over="ignore"):
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        x += s
    return x.T.copy()


def _keys_for(seed, start, count):
    base = int.from_bytes(bytes(seed), "big")
    keys = np.empty((count, 8), dtype=np.uint32)
    for t in range(count):
        kb = ((base + start + t) % (1 << 256)).to_bytes(32, "big")
        keys[t] = np.frombuffer(kb, dtype="<u4")
    return keys


def crs_coeffs(seed, start, count):
    """structs.rs:147-171 for `count` consecutive counters starting at base_seed + start."""
    keys = _keys_for(seed, start, count)
    out = np.empty(count, dtype=np.uint32)
    zone = (Q << 115) - 1
    pending = list(range(count))
    word = {t: 0 for t in pending}
    cache = {}
    while pending:
        # all pending coefficients read 4 keystream words at their own position
        blocks_needed = sorted({(word[t] + u) // 16 for t in pending for u in range(4)})
        for bk in blocks_needed:
            if bk not in cache:
                cache[bk] = chacha20_blocks(keys, bk)
        nxt = []
        for t in pending:
            w = [int(cache[(word[t] + u) // 16][t, (word[t] + u) % 16]) for u in range(4)]
            word[t] += 4
            v = w[0] | (w[1] << 32) | (w[2] << 64) | (w[3] << 96)
            prod = v * Q
            hi, lo = prod >> 128, prod & ((1 << 128) - 1)
            if lo <= zone:
                out[t] = hi
            else:
                nxt.append(t)
        pending = nxt
    return out


def crs_polys(seed, start, n):
    return crs_coeffs(seed, start, n * D).reshape(n, D)


def rq_mul(a, b):
    full = np.convolve(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64))
    full = np.concatenate([full, np.zeros(2 * D - full.size, dtype=np.int64)])
    return ((full[:D] - full[D:]) % Q).astype(np.uint32)


def inner(v1, v2):
    acc = np.zeros(D, dtype=np.int64)
    for x, y in zip(np.asarray(v1).reshape(-1, D), np.asarray(v2).reshape(-1, D)):
        acc += rq_mul(x, y)
    return (acc % Q).astype(np.uint32)


def sigma_inv(a):
    a = np.asarray(a, dtype=np.int64)
    out = np.zeros(D, dtype=np.int64)
    out[0] = a[0]
    out[1:] = (-a[:0:-1]) % Q
    return out.astype(np.uint32)


def decompose(p, base, exp):
    v = np.asarray(p, dtype=np.int64).copy()
    out = np.zeros((exp, D), dtype=np.uint32)
    for k in range(exp):
        x = v % base
        out[k] = np.where(x <= base // 2, x, base - x)
        v //= base
    return out


class _Zq:
    """Operator semantics of algebraic.rs:24-297 needed by decompose_polynomial."""

    def __init__(self, v):
        self.v = v % Q

    def gt(self, rhs):          # PartialOrd<i128>
        return self.v > rhs % Q

    def ne0(self):
        return self.v != 0


def _trunc_rem(a, b):           # Rust % on i128 (sign of dividend); operands here are >= 0
    return a - b * int(a / b) if a < 0 else a % b


def decompose_literal(p, base, exp):
    out = np.zeros((exp, D), dtype=np.uint32)
    for deg in range(D):
        coeff = _Zq(int(p[deg]))
        k = 0
        while coeff.ne0():
            r0 = _Zq(_trunc_rem(coeff.v, base))
            rem = _Zq(_trunc_rem(base - r0.v, base)) if r0.gt(base // 2) else r0
            if k < exp:
                out[k, deg] = rem.v
            k += 1
            coeff = _Zq(_Zq(coeff.v - rem.v).v // base)
    return out


def offsets(N, R, T_1):
    """structs.rs:55-144 offset arithmetic as Python big ints (kappa = kappa1 = kappa2 = N*D)."""
    K = N * D
    offA = K * N * D
    sizeB = K * K
    endB = offA + R * T_1 * sizeB * D

    def sp(i):
        return i * R - i * (i - 1) // 2 if i > 0 else 0

    return {
        "A": lambda row: row * N * D,
        "B": lambda i, k, row: offA + (i * T_1 + k) * sizeB + row * K * D,
        "C": lambda i, j, k: endB + (k + T_1 * (sp(i) + (j - i))) * K * D,
        "D": lambda i, j, k: endB + (R * (R + 1) // 2) * K * D + (k + T_1 * (sp(i) + (j - i))) * K * D,
    }


def prove(N, R, consts, seed, S, phi, a, b, ch):
    """proofgen.rs:30-427 with K = L = 1. consts: dict with B_1,T_1,B_2,T_2,BETA_BOUND."""
    K = N * D
    T1, B1, T2, B2 = consts["T_1"], consts["B_1"], consts["T_2"], consts["B_2"]
    off = offsets(N, R, T1)
    S = np.asarray(S).reshape(R, N, D)
    t = np.zeros((R, K, D), dtype=np.uint32)
    for row in range(K):
        A = crs_polys(seed, off["A"](row), N)
        for i in range(R):
            t[i, row] = inner(A, S[i])
    g = np.stack([[inner(S[i], S[j]) for j in range(R)] for i in range(R)])
    u1 = np.zeros((K, D), dtype=np.int64)
    tdec = np.stack([np.stack([decompose(t[i, y], B1, T1) for y in range(K)], axis=1) for i in range(R)])  # [R][T1][K][D]
    for x in range(K):
        for i in range(R):
            for k in range(T1):
                u1[x] += inner(crs_polys(seed, off["B"](i, k, x), K), tdec[i, k])
    for i in range(R):
        for j in range(i, R):
            gd = decompose(g[i, j], B2, T2)
            for k in range(T2):
                Cv = crs_polys(seed, off["C"](i, j, k), K)
                for x in range(K):
                    u1[x] += rq_mul(gd[k], Cv[x])
    u1 = (u1 % Q).astype(np.uint32)
    # JL
    att = 0
    while True:
        pi = np.asarray(ch["pi"][att], dtype=np.int64)           # [R][256][N*D]
        p = np.einsum("ijc,ic->j", pi, S.reshape(R, N * D).astype(np.int64))
        if np.sqrt(float(int((p.astype(object) ** 2).sum()))) <= np.sqrt(128.0) * float(consts["BETA_BOUND"]):
            break
        att += 1
        if att > 5 or att >= len(ch["pi"]):
            raise RuntimeError("failed JL")
    psi, omega = int(ch["psi"]), np.asarray(ch["omega"], dtype=np.int64)
    phi = np.asarray(phi).reshape(R, N, D)
    v = np.einsum("j,ijc->ic", omega, pi % Q) % Q                # Pi^T omega, [R][N*D]
    phipp = np.zeros((R, N, D), dtype=np.uint32)
    for i in range(R):
        for n in range(N):
            phipp[i, n] = (psi * phi[i, n].astype(np.int64) + sigma_inv(v[i, n * D:(n + 1) * D])) % Q
    a = np.asarray(a).reshape(R, R, D)
    bpp = np.zeros(D, dtype=np.int64)
    for i in range(R):
        for j in range(R):
            bpp += rq_mul((psi * a[i, j].astype(np.int64)) % Q, g[i, j])
        bpp += inner(phipp[i], S[i])
    bpp = (bpp % Q).astype(np.uint32)
    proj = (p % Q).astype(np.uint32)
    assert int(bpp[0]) == (int((omega * proj).sum()) + psi * int(b[0])) % Q, "verify_b_prime_prime"
    alpha, beta = ch["alpha"], ch["beta"]
    phif = np.zeros((R, N, D), dtype=np.uint32)
    for i in range(R):
        for n in range(N):
            phif[i, n] = (rq_mul(alpha, phi[i, n]).astype(np.int64) + rq_mul(beta, phipp[i, n])) % Q
    h = np.zeros((R, R, D), dtype=np.uint32)
    for i in range(R):
        for j in range(R):
            h[i, j] = ((inner(phif[i], S[j]).astype(np.int64) + inner(phif[j], S[i])) * 4096) % Q
    u2 = np.zeros((K, D), dtype=np.int64)
    for i in range(R):
        for j in range(i, R):
            hd = decompose(h[i, j], B1, T1)
            for k in range(T1):
                Dv = crs_polys(seed, off["D"](i, j, k), K)
                for x in range(K):
                    u2[x] += rq_mul(hd[k], Dv[x])
    u2 = (u2 % Q).astype(np.uint32)
    z = np.zeros((N, D), dtype=np.int64)
    for i in range(R):
        for n in range(N):
            z[n] += rq_mul(ch["c"][i], S[i, n])
    z = (z % Q).astype(np.uint32)
    return {"t": t, "g": g, "u_1": u1, "projection_int": p.astype(np.int64), "projection": proj,
            "b_prime_prime": bpp, "phi_final": phif, "h": h, "u_2": u2, "z": z, "jl_attempt": att}
