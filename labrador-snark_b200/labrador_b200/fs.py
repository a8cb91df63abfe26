"""Fiat-Shamir seed chain of liblabrador_b200 (include/labrador_b200.h, lab_fs_*), restated with hashlib.

The reference simulates the interactive protocol in-process and lists Fiat-Shamir as TODO (README.md:12); this library derives
the verifier's randomness from the transcript prefix.  This module is the independent restatement the parity tests use: the
challenges it derives from a GPU transcript must equal the ones lab_prove_fs returned, and the CPU oracle, given those
challenges by injection, must reproduce the transcript.  Order of consumption: SURVEY appendix A.1."""
import hashlib
import struct

import numpy as np

from . import synth

DOMAIN = b"LaBRADOR-B200-FS-v1"


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32).tobytes()


def init(N, R, crs_seed, phi, a, b):
    return hashlib.sha256(DOMAIN + bytes(crs_seed) + struct.pack("<QQ", N, R) + _u32(phi) + _u32(a) + _u32(b)).digest()


def absorb(state, label, data):
    return hashlib.sha256(state + label.encode() + data).digest()


def squeeze(state, label, index=0):
    return int.from_bytes(hashlib.sha256(state + label.encode() + struct.pack("<I", index)).digest()[:8], "little")


def derive_challenges(N, R, crs_seed, phi, a, b, tr):
    """tr: transcript dict (u_1, jl_attempt, projection_int, b_prime_prime, u_2).  Returns the challenges dict the derived verifier
    answers with: pi holds attempts 0 .. jl_attempt (so that jl_attempt indexes it like in the interactive runs)."""
    st = init(N, R, crs_seed, phi, a, b)
    st = absorb(st, "u_1", _u32(tr["u_1"]))
    att = int(tr["jl_attempt"])
    pi = np.stack([synth.sample_pi(N, R, squeeze(st, "pi", t), 0) for t in range(att + 1)])
    st = absorb(st, "proj", struct.pack("<I", att) + np.ascontiguousarray(tr["projection_int"], dtype=np.int64).tobytes())
    s_agg = squeeze(st, "agg")
    psi, omega = int(synth.prg_zq(s_agg, 6, 1)[0]), synth.prg_zq(s_agg, 7, synth.JL)
    st = absorb(st, "bpp", _u32(tr["b_prime_prime"]))
    s_ab = squeeze(st, "ab")
    alpha, beta = synth.prg_zq(s_ab, 8, synth.D), synth.prg_zq(s_ab, 9, synth.D)
    st = absorb(st, "u_2", _u32(tr["u_2"]))
    s_c = squeeze(st, "c")
    c = np.stack([synth.sample_challenge_poly(s_c, i) for i in range(R)])
    return {"pi": pi, "psi": psi, "omega": omega, "alpha": alpha, "beta": beta, "c": c}
