#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json from the oracle.

The reference itself publishes no golden vectors and cannot be run here (Rust, no toolchain), so
these fixtures pin the ORACLE's current behaviour: external known answers (RFC 7539 / rand_chacha ChaCha20
vectors, the constants table of SURVEY F8, the CRS coefficients SURVEY 8c derived independently) are asserted
in tests/test_oracle_kat.py; this file adds digests of oracle outputs so that any drift is caught."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    seed = bytes(range(32))
    out = {"crs_seed": seed.hex()}
    out["crs_first_poly_seed00"] = oracle.crs_poly(bytes(32), 0).tolist()
    out["crs_poly_seed_00_1f_ctr_2^64+5"] = oracle.crs_poly(seed, 2**64 + 5).tolist()
    out["constants"] = {}
    for (N, R) in [(1, 1), (1, 2), (2, 2), (2, 4), (4, 4), (4, 8), (8, 8), (8, 16), (16, 16), (16, 32), (32, 32), (4096, 64)]:
        c, rc = oracle.constants(N, R)
        out["constants"][f"{N},{R}"] = {"rc": rc, "BETA_BOUND": c.BETA_BOUND, "B": c.B, "T_1": c.T_1 if rc == 0 else None, "B_1": c.B_1,
                                        "T_2": c.T_2 if rc == 0 else None, "B_2": c.B_2, "BETA_PRIME": c.BETA_PRIME if rc == 0 else None}
    proofs = {}
    for (N, R, s) in [(1, 1, 11), (1, 2, 12), (2, 2, 13)]:
        c, _ = oracle.constants(N, R)
        S = oracle.generate_witness(c, s)
        phi, a, b = oracle.generate_state(c, S, s)
        ch = oracle.sample_challenges(c, s, 2)
        rc, tr = oracle.prove(c, seed, S, phi, a, b, ch, ntt=False, nthreads=8)
        ok, failed, norm = oracle.verify(c, seed, phi, a, b, ch, tr, ntt=False, nthreads=8)
        proofs[f"{N},{R},{s}"] = {"rc": rc, "verify": ok, "norm_sum": norm, "jl_attempt": tr["jl_attempt"],
                                  "digest": {k: digest(tr[k]) for k in ("t", "g", "u_1", "projection_int", "b_prime_prime", "phi_final", "h", "u_2", "z")},
                                  "witness": digest(S), "z0": tr["z"][0, :8].tolist(), "u1_0": tr["u_1"][0, :8].tolist()}
    out["proofs"] = proofs
    # wire format: sha256 and length of bincode::serialize(&Transcript) (structs.rs:192-221) for the (1,2) proof above, produced by
    # the independent struct.pack restatement in tests/test_host.py (the product's lab_transcript_bincode must reproduce it)
    sys.path.insert(0, os.path.join(HERE, ".."))
    sys.path.insert(0, os.path.join(HERE, "..", "..", "labrador-snark_b200"))
    from test_host import _bincode_reference
    import labrador_b200 as lb
    c, _ = oracle.constants(1, 2)
    S = oracle.generate_witness(c, 12)
    phi, a, b = oracle.generate_state(c, S, 12)
    ch = oracle.sample_challenges(c, 12, 2)
    rc, tr = oracle.prove(c, seed, S, phi, a, b, ch, ntt=False, nthreads=8)
    blob = _bincode_reference(lb.RuntimeConstants.new(1, 2), tr, ch, int(tr["jl_attempt"]))
    out["bincode_1,2,12"] = {"len": len(blob), "sha256": hashlib.sha256(blob).hexdigest()}
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
