#!/usr/bin/env python3
"""Writes tests/golden/prove_8_8.json: digests of the ORACLE's transcript of one seeded (N, R) = (8, 8) statement (T_1 = 5).

bench.py proves the same statement on the GPU -- row- and vector-sharded over the ranks by the library's own communicator --
and compares the transcript digest of every rank with this file (extra.sharded_prove.matches_oracle), so that the scaling
runs carry a proof that lab_comm_* changes no bit.  Inputs: synth.generate_witness / generate_state / sample_challenges with
PRG seed 0x4C61425241444F52 + 88 (the oracle's generators; tests/test_host.py checks that labrador_b200.synth equals them),
CRS seed 00..1f, three JL attempts."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle  # noqa: E402

FIELDS = ("t", "g", "u_1", "projection_int", "b_prime_prime", "h", "u_2", "z")
N, R, SEED, ATTEMPTS = 8, 8, 0x4C61425241444F52 + 88, 3


def digest(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    seed32 = bytes(range(32))
    c, rc = oracle.constants(N, R)
    assert rc == 0
    S = oracle.generate_witness(c, SEED)
    phi, a, b = oracle.generate_state(c, S, SEED)
    ch = oracle.sample_challenges(c, SEED, ATTEMPTS)
    t0 = time.time()
    rc, tr = oracle.prove(c, seed32, S, phi, a, b, ch, ntt=True, nthreads=oracle.num_threads())
    assert rc == 0
    ok, failed, norm = oracle.verify(c, seed32, phi, a, b, ch, tr, ntt=True, nthreads=oracle.num_threads())
    assert ok
    out = {"N": N, "R": R, "prg_seed": SEED, "crs_seed": seed32.hex(), "attempts": ATTEMPTS, "jl_attempt": int(tr["jl_attempt"]), "norm_sum": int(norm),
           "inputs_sha256": digest([S, phi, a, b, ch["pi"], np.array([ch["psi"]], np.uint32), ch["omega"], ch["alpha"], ch["beta"], ch["c"]]),
           "transcript_sha256": digest([tr[k] for k in FIELDS]), "fields": list(FIELDS),
           "per_field": {k: digest([tr[k]]) for k in FIELDS}, "oracle_seconds": round(time.time() - t0, 1)}
    json.dump(out, open(os.path.join(HERE, "prove_8_8.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
