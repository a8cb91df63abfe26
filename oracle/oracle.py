"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- never imported by the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/labrador_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblabrador_oracle.so")

D = 64
Q = 8191
JL = 256
SEED = 0x4C61425241444F52


class Constants(C.Structure):
    _fields_ = [
        ("N", C.c_uint64), ("R", C.c_uint64), ("BETA_BOUND", C.c_int64), ("STD", C.c_double),
        ("B", C.c_int64), ("T_1", C.c_int64), ("B_1", C.c_int64), ("T_2", C.c_int64), ("B_2", C.c_int64),
        ("GAMMA", C.c_double), ("GAMMA_1", C.c_double), ("GAMMA_2", C.c_double), ("BETA_PRIME", C.c_double),
        ("KAPPA", C.c_uint64), ("KAPPA_1", C.c_uint64), ("KAPPA_2", C.c_uint64), ("degenerate", C.c_int),
    ]


class State(C.Structure):
    _fields_ = [("phi", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p)]


class Challenges(C.Structure):
    _fields_ = [("pi", C.c_void_p), ("n_attempts", C.c_int), ("psi", C.c_uint32),
                ("omega", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p), ("c", C.c_void_p)]


class Transcript(C.Structure):
    _fields_ = [("u_1", C.c_void_p), ("jl_attempt", C.c_int), ("projection_int", C.c_void_p),
                ("projection", C.c_void_p), ("b_prime_prime", C.c_void_p), ("u_2", C.c_void_p),
                ("z", C.c_void_p), ("t", C.c_void_p), ("g", C.c_void_p), ("h", C.c_void_p),
                ("phi_final", C.c_void_p)]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "labrador_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.lo_prg_u64.restype = C.c_uint64
        _lib.lo_prg_u64.argtypes = [C.c_uint64] * 3
        _lib.lo_prg_zq.restype = C.c_uint32
        _lib.lo_prg_zq.argtypes = [C.c_uint64] * 3
        _lib.lo_norm_sq.restype = C.c_uint64
        _lib.lo_crs_coeff.restype = C.c_uint32
        _lib.lo_mod_positive.restype = C.c_uint32
        _lib.lo_mod_positive.argtypes = [C.c_int64]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def constants(N, R):
    c = Constants()
    rc = lib().lo_runtime_constants(C.c_uint64(N), C.c_uint64(R), C.byref(c))
    return c, rc


def rq_mul(a, b, ntt=False):
    a, b = _u32(a), _u32(b)
    out = np.empty(D, np.uint32)
    (lib().lo_rq_mul_ntt if ntt else lib().lo_rq_mul)(_p(a), _p(b), _p(out))
    return out


def rq_mul_batch(a, b, ntt=False):
    a, b = _u32(a).reshape(-1, D), _u32(b).reshape(-1, D)
    out = np.empty_like(a)
    f = lib().lo_rq_mul_ntt if ntt else lib().lo_rq_mul
    for i in range(a.shape[0]):
        f(_p(a[i]), _p(b[i]), _p(out[i]))
    return out


def inner_product(v1, v2):
    v1, v2 = _u32(v1).reshape(-1, D), _u32(v2).reshape(-1, D)
    out = np.empty(D, np.uint32)
    lib().lo_inner_product(_p(v1), _p(v2), C.c_size_t(v1.shape[0]), _p(out))
    return out


def sigma_inv(a):
    a = _u32(a)
    out = np.empty(D, np.uint32)
    lib().lo_sigma_inv(_p(a), _p(out))
    return out


def decompose(p, base, exp, literal=False):
    p = _u32(p)
    out = np.empty((exp, D), np.uint32)
    if literal:
        rc = lib().lo_decompose_literal(_p(p), C.c_int64(base), C.c_int64(exp), _p(out))
        assert rc == 0
    else:
        lib().lo_decompose(_p(p), C.c_int64(base), C.c_int64(exp), _p(out))
    return out


def norm_sq(x):
    x = _u32(x).reshape(-1)
    return int(lib().lo_norm_sq(_p(x), C.c_size_t(x.size)))


def ntt_fwd(poly):
    poly = _u32(poly)
    out = np.empty(D, np.uint32)
    lib().lo_ntt_fwd(_p(poly), _p(out))
    return out


def ntt_inv(re_im):
    re_im = _u32(re_im)
    out = np.empty(D, np.uint32)
    lib().lo_ntt_inv(_p(re_im), _p(out))
    return out


def ntt_slot_exponents():
    return [lib().lo_ntt_slot_exponent(j) for j in range(32)]


def chacha20_block(key_words, counter=0, stream=0):
    key = _u32(key_words)
    out = np.empty(16, np.uint32)
    lib().lo_chacha20_block(_p(key), C.c_uint64(counter), C.c_uint64(stream), _p(out))
    return out


def _seed(seed):
    s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    assert s.size == 32
    return s


def crs_poly(seed, start):
    """64 coefficients starting at absolute counter offset `start` (python int < 2^128)."""
    s = _seed(seed)
    out = np.empty(D, np.uint32)
    # unsigned __int128 by value: SysV passes it as two 64-bit integer registers (lo, hi)
    lib().lo_crs_poly(_p(s), C.c_uint64(start & (2**64 - 1)), C.c_uint64(start >> 64), _p(out))
    return out


def crs_polys(seed, start, n):
    return np.stack([crs_poly(seed, start + D * i) for i in range(n)])


def offset(which, c, i=0, j=0, k=0, row=0):
    lo, hi = C.c_uint64(), C.c_uint64()
    w = {"A": 0, "B": 1, "C": 2, "D": 3}[which]
    lib().lo_off_split(w, C.byref(c), C.c_uint64(i), C.c_uint64(j), C.c_uint64(k), C.c_uint64(row), C.byref(lo), C.byref(hi))
    return lo.value | (hi.value << 64)


def fetch_A_row(c, seed, row):
    s = _seed(seed)
    out = np.empty((c.N, D), np.uint32)
    lib().lo_fetch_A_row(C.byref(c), _p(s), C.c_uint64(row), _p(out))
    return out


def fetch_B_ik_row(c, seed, i, k, row):
    s = _seed(seed)
    out = np.empty((c.KAPPA, D), np.uint32)
    lib().lo_fetch_B_ik_row(C.byref(c), _p(s), C.c_uint64(i), C.c_uint64(k), C.c_uint64(row), _p(out))
    return out


def fetch_C_ijk(c, seed, i, j, k):
    s = _seed(seed)
    out = np.empty((c.KAPPA_2, D), np.uint32)
    lib().lo_fetch_C_ijk(C.byref(c), _p(s), C.c_uint64(i), C.c_uint64(j), C.c_uint64(k), _p(out))
    return out


def fetch_D_ijk(c, seed, i, j, k):
    s = _seed(seed)
    out = np.empty((c.KAPPA_2, D), np.uint32)
    lib().lo_fetch_D_ijk(C.byref(c), _p(s), C.c_uint64(i), C.c_uint64(j), C.c_uint64(k), _p(out))
    return out


def commit_inner_rows(c, seed, S, row0, nrows, ntt=False, nthreads=1):
    s = _seed(seed)
    S = _u32(S)
    T = np.empty((c.R, nrows, D), np.uint32)
    lib().lo_commit_inner_rows(C.byref(c), _p(s), _p(S), C.c_uint64(row0), C.c_uint64(nrows), C.c_int(int(ntt)), C.c_int(nthreads), _p(T))
    return T


def gram(c, S):
    S = _u32(S)
    G = np.empty((c.R, c.R, D), np.uint32)
    lib().lo_gram(C.byref(c), _p(S), _p(G))
    return G


def jl_project(c, S, Pi):
    S = _u32(S)
    Pi = np.ascontiguousarray(Pi, dtype=np.int8)
    p = np.empty(JL, np.int64)
    lib().lo_jl_project(C.byref(c), _p(S), _p(Pi), _p(p))
    return p


def valid_projection(c, p):
    p = np.ascontiguousarray(p, dtype=np.int64)
    return bool(lib().lo_valid_projection(C.byref(c), _p(p)))


def amortize_z(c, S, ch):
    S, ch = _u32(S), _u32(ch)
    z = np.empty((c.N, D), np.uint32)
    lib().lo_amortize_z(C.byref(c), _p(S), _p(ch), _p(z))
    return z


def generate_witness(c, seed=SEED):
    S = np.empty((c.R, c.N, D), np.uint32)
    lib().lo_generate_witness(C.byref(c), C.c_uint64(seed), _p(S))
    return S


def generate_state(c, S, seed=SEED):
    S = _u32(S)
    phi = np.empty((c.R, c.N, D), np.uint32)
    a = np.empty((c.R, c.R, D), np.uint32)
    b = np.empty(D, np.uint32)
    lib().lo_generate_state(C.byref(c), C.c_uint64(seed), _p(S), _p(phi), _p(a), _p(b))
    return phi, a, b


def sample_pi(c, seed=SEED, attempt=0):
    pi = np.empty((c.R, JL, c.N * D), np.int8)
    lib().lo_sample_pi(C.byref(c), C.c_uint64(seed), C.c_uint64(attempt), _p(pi))
    return pi


def sample_challenge_poly(seed, idx):
    out = np.empty(D, np.uint32)
    lib().lo_sample_challenge_poly(C.c_uint64(seed), C.c_uint64(idx), _p(out))
    return out


def prg_zq(seed, stream, n):
    return np.array([lib().lo_prg_zq(seed, stream, i) for i in range(n)], dtype=np.uint32)


def sample_challenges(c, seed=SEED, n_attempts=1):
    """Verifier randomness in the reference's consumption order, seeded."""
    pi = np.stack([sample_pi(c, seed, a) for a in range(n_attempts)])
    return {
        "pi": pi,
        "psi": int(lib().lo_prg_zq(seed, 6, 0)),
        "omega": prg_zq(seed, 7, JL),
        "alpha": prg_zq(seed, 8, D),
        "beta": prg_zq(seed, 9, D),
        "c": np.stack([sample_challenge_poly(seed, i) for i in range(c.R)]),
    }


def _mk_state(phi, a, b):
    keep = [_u32(phi), _u32(a), _u32(b)]
    st = State(_p(keep[0]), _p(keep[1]), _p(keep[2]))
    return st, keep


def _mk_chal(ch):
    keep = [np.ascontiguousarray(ch["pi"], dtype=np.int8), _u32(ch["omega"]), _u32(ch["alpha"]), _u32(ch["beta"]), _u32(ch["c"])]
    n_att = keep[0].shape[0] if keep[0].ndim == 4 else 1
    cc = Challenges(_p(keep[0]), n_att, int(ch["psi"]), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]))
    return cc, keep


def _alloc_transcript(c):
    bufs = {
        "u_1": np.zeros((c.KAPPA_1, D), np.uint32),
        "projection_int": np.zeros(JL, np.int64),
        "projection": np.zeros(JL, np.uint32),
        "b_prime_prime": np.zeros(D, np.uint32),
        "u_2": np.zeros((c.KAPPA_2, D), np.uint32),
        "z": np.zeros((c.N, D), np.uint32),
        "t": np.zeros((c.R, c.KAPPA, D), np.uint32),
        "g": np.zeros((c.R, c.R, D), np.uint32),
        "h": np.zeros((c.R, c.R, D), np.uint32),
        "phi_final": np.zeros((c.R, c.N, D), np.uint32),
    }
    return bufs


def _mk_transcript(bufs, jl_attempt=0):
    return Transcript(_p(bufs["u_1"]), jl_attempt, _p(bufs["projection_int"]), _p(bufs["projection"]),
                      _p(bufs["b_prime_prime"]), _p(bufs["u_2"]), _p(bufs["z"]), _p(bufs["t"]),
                      _p(bufs["g"]), _p(bufs["h"]), _p(bufs["phi_final"]))


def prove(c, seed, S, phi, a, b, ch, ntt=False, nthreads=1):
    """Returns (status, transcript dict)."""
    s = _seed(seed)
    S = _u32(S)
    st, k1 = _mk_state(phi, a, b)
    cc, k2 = _mk_chal(ch)
    bufs = _alloc_transcript(c)
    tr = _mk_transcript(bufs)
    rc = lib().lo_prove(C.byref(c), _p(s), _p(S), C.byref(st), C.byref(cc), C.c_int(int(ntt)), C.c_int(nthreads), C.byref(tr))
    bufs["jl_attempt"] = tr.jl_attempt
    del k1, k2
    return rc, bufs


def verify(c, seed, phi, a, b, ch, transcript, ntt=False, nthreads=1):
    """Returns (accept: bool, failed_check: int, norm_sum: int)."""
    s = _seed(seed)
    st, k1 = _mk_state(phi, a, b)
    cc, k2 = _mk_chal(ch)
    bufs = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in transcript.items()}
    for k in ("u_1", "projection", "b_prime_prime", "u_2", "z", "t", "g", "h", "phi_final"):
        bufs[k] = _u32(bufs[k])
    bufs["projection_int"] = np.ascontiguousarray(bufs["projection_int"], dtype=np.int64)
    tr = _mk_transcript(bufs, int(transcript.get("jl_attempt", 0)))
    fc = C.c_int(0)
    ns = C.c_uint64(0)
    ok = lib().lo_verify(C.byref(c), _p(s), C.byref(st), C.byref(cc), C.byref(tr), C.c_int(int(ntt)), C.c_int(nthreads), C.byref(fc), C.byref(ns))
    del k1, k2
    return bool(ok), fc.value, ns.value


def num_threads():
    return lib().lo_num_threads_default()
