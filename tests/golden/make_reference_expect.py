#!/usr/bin/env python3
"""Writes tests/golden/reference_expect.json: what THIS repo's CPU oracle (and its byte-exact bincode writer) says the unmodified
reference crate must output for labrador-snark_b200/rust/reference-vectors/examples/gen_vectors.rs.  The seed-dependent part of
that generator (CRS offsets under the reference's private random base seed) is recomputed at test time from the seed the
generator reports.  See tests/test_reference_vectors.py."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, "..", "..", "labrador-snark_b200"))
import oracle  # noqa: E402
import labrador_b200 as lb  # noqa: E402

Q, D = 8191, 64


def fnv(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def poly(tag, idx):
    """The fixed polynomials of gen_vectors.rs: coefficient d of polynomial (tag, idx)."""
    out = np.zeros(D, np.uint32)
    if idx == 3:
        return out
    ln = 64 - (idx % 5)
    for d in range(ln):
        out[d] = (tag * 131 + idx * 17 + d * 7 + 1) % Q
    return out


def perf_shapes():
    n, r, out = 1, 2, []
    for size_pow in range(2, 11):
        if size_pow % 2 == 0:
            n *= 2
        else:
            r *= 2
        out.append((n, r))
    return out


def fixed_transcript():
    N, R, kappa, nd = 1, 2, 64, 64
    polys = lambda tag, cnt: np.stack([poly(tag, i) for i in range(cnt)])
    pi = np.zeros((1, R, 256, nd), np.int8)
    for i in range(R):
        for j in range(256):
            for x in range(nd):
                v = (i * 7 + j * 3 + x * 5) % 4
                pi[0, i, j, x] = -1 if v == 0 else (1 if v == 3 else 0)
    tr = {"u_1": polys(1, kappa), "projection": np.array([(j * 37 + 5) % Q for j in range(256)], np.uint32), "b_prime_prime": poly(5, 0),
          "u_2": polys(8, kappa), "z": polys(10, 1), "t": np.stack([np.stack([poly(11, i * kappa + k) for k in range(kappa)]) for i in range(R)]),
          "g": np.stack([poly(12, e) for e in range(R * R)]).reshape(R, R, D), "h": np.stack([poly(13, e) for e in range(R * R)]).reshape(R, R, D),
          "jl_attempt": 0}
    ch = {"pi": pi, "psi": 123, "omega": np.array([(j * 11 + 3) % Q for j in range(256)], np.uint32), "alpha": poly(6, 0), "beta": poly(7, 0),
          "c": polys(9, R)}
    return lb.RuntimeConstants.new(N, R), tr, ch


def expected():
    s1 = bytes(range(32))
    s2 = bytes([0x7F]) + bytes([0xFF]) * 31
    out = {"random_oracle_gen": {"seed_zero": oracle.crs_poly(bytes(32), 0).tolist(), "seed_00_1f": oracle.crs_poly(s1, 0).tolist(),
                                 "seed_7fff_ff": oracle.crs_poly(s2, 0).tolist()}}
    cs = {}
    for (n, r) in perf_shapes():
        c, rc = oracle.constants(n, r)
        cs[f"{n},{r}"] = {"BETA_BOUND": c.BETA_BOUND, "STD": c.STD, "B": c.B, "T_1": c.T_1, "B_1": c.B_1, "T_2": c.T_2, "B_2": c.B_2, "GAMMA": c.GAMMA,
                          "GAMMA_1": c.GAMMA_1, "GAMMA_2": c.GAMMA_2, "BETA_PRIME": c.BETA_PRIME, "KAPPA": c.KAPPA}
    out["constants"] = cs
    x, y = poly(21, 1), poly(22, 2)
    out["rq_mul"] = {"schoolbook": oracle.rq_mul(x, y, ntt=False).tolist(), "ntt": oracle.rq_mul(x, y, ntt=True).tolist()}
    c, tr, ch = fixed_transcript()
    raw = lb.api.transcript_bincode(c, tr, ch)
    gz, n = lb.api.transcript_size_in_bytes(c, tr, ch)
    out["bincode"] = {"len": len(raw), "fnv": fnv(raw), "head": raw[:64].hex(), "tail": raw[-64:].hex(), "size_in_bytes": gz,
                      "size_in_bytes_note": "zlib level 9; flate2's miniz_oxide backend may differ by a few bytes (a size metric, not byte parity)"}
    return out


def crs_expected(base_seed_hex, N=2, R=2):
    """The seed-dependent values of gen_vectors.rs section 2, recomputed by the oracle for the seed the reference drew."""
    seed = bytes.fromhex(base_seed_hex)
    c, _ = oracle.constants(N, R)
    dense = lambda polys: np.asarray(polys, dtype=np.uint32)
    h = lambda polys: fnv(dense(polys).astype("<u2").tobytes())
    b113, c011, d112 = oracle.fetch_B_ik_row(c, seed, 1, 1, 3), oracle.fetch_C_ijk(c, seed, 0, 1, 1), oracle.fetch_D_ijk(c, seed, 1, 1, 2)
    return {"A_row_5": oracle.fetch_A_row(c, seed, 5).tolist(), "B_1_1_3_first2": b113[:2].tolist(), "B_1_1_3_fnv": h(b113),
            "C_0_1_1_first2": c011[:2].tolist(), "C_0_1_1_fnv": h(c011), "D_1_1_2_first2": d112[:2].tolist(), "D_1_1_2_fnv": h(d112),
            "B_0_1_0_equals_B_0_0_2": bool(np.array_equal(oracle.fetch_B_ik_row(c, seed, 0, 1, 0)[0], oracle.fetch_B_ik_row(c, seed, 0, 0, 2)[0]))}


if __name__ == "__main__":
    json.dump(expected(), open(os.path.join(HERE, "reference_expect.json"), "w"), indent=1)
    print("wrote reference_expect.json")
