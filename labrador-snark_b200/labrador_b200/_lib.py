"""Loads liblabrador_b200.so (the CUDA product) and declares the C ABI of include/labrador_b200.h.

There is no fallback of any kind: if the shared library is missing or no CUDA device is present the
import / context creation raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "..", "liblabrador_b200.so")

D = 64
Q = 8191
JL_ROWS = 256

STATUS = {0: "LAB_OK", 1: "LAB_ERR_JL_REJECTED", 2: "LAB_ERR_BPP_CHECK", 3: "LAB_ERR_SHAPE",
          4: "LAB_ERR_PARAMS", 5: "LAB_ERR_CUDA", 6: "LAB_ERR_ALLOC"}


class LabError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS.get(status, status)}: {msg}")
        self.status = status


class Constants(C.Structure):
    """RuntimeConstants (constants.rs:205-265)."""
    _fields_ = [
        ("N", C.c_uint64), ("R", C.c_uint64), ("BETA_BOUND", C.c_int64), ("STD", C.c_double),
        ("B", C.c_int64), ("T_1", C.c_int64), ("B_1", C.c_int64), ("T_2", C.c_int64), ("B_2", C.c_int64),
        ("GAMMA", C.c_double), ("GAMMA_1", C.c_double), ("GAMMA_2", C.c_double), ("BETA_PRIME", C.c_double),
        ("KAPPA", C.c_uint64), ("KAPPA_1", C.c_uint64), ("KAPPA_2", C.c_uint64), ("degenerate", C.c_int),
    ]


class CState(C.Structure):
    _fields_ = [("phi", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p)]


class CChallenges(C.Structure):
    _fields_ = [("pi", C.c_void_p), ("n_attempts", C.c_int), ("psi", C.c_uint32),
                ("omega", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p), ("c", C.c_void_p),
                ("pi2", C.c_void_p)]          # optional 2-bit packed twin of pi (include/labrador_b200.h)


class CChallengesBuf(C.Structure):
    """lab_challenges_buf: writable challenges of one (the accepted) JL attempt."""
    _fields_ = [("pi2", C.c_void_p), ("psi", C.c_uint32), ("omega", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p), ("c", C.c_void_p)]


class CTranscript(C.Structure):
    _fields_ = [("u_1", C.c_void_p), ("jl_attempt", C.c_int), ("projection_int", C.c_void_p),
                ("projection", C.c_void_p), ("b_prime_prime", C.c_void_p), ("u_2", C.c_void_p),
                ("z", C.c_void_p), ("t", C.c_void_p), ("g", C.c_void_p), ("h", C.c_void_p),
                ("phi_final", C.c_void_p), ("norm_sum", C.c_uint64)]


# every symbol include/labrador_b200.h declares (tests assert the .so exports all of them)
SYMBOLS = [
    "lab_ctx_create", "lab_ctx_destroy", "lab_last_error", "lab_sync", "lab_malloc", "lab_free",
    "lab_memcpy_h2d", "lab_memcpy_d2h", "lab_stream", "lab_kernel_launches", "lab_version", "lab_timer_start", "lab_timer_stop",
    "lab_runtime_constants", "lab_ntt_fwd_batch", "lab_ntt_inv_batch", "lab_polymul_batch",
    "lab_ntt_fwd_batch_dev", "lab_ntt_inv_batch_dev", "lab_polymul_batch_dev", "lab_ntt_slot_exponents",
    "lab_inner_product_batch", "lab_decompose", "lab_norm_sq", "lab_norm_sq_dev", "lab_sigma_inv",
    "lab_crs_expand", "lab_crs_expand_dev", "lab_crs_fetch", "lab_crs_offset",
    "lab_commit_inner", "lab_gram", "lab_jl_project", "lab_jl_project_part", "lab_commit_outer_u1", "lab_commit_outer_u2",
    "lab_aggregate_phi", "lab_h_gram", "lab_amortize_z", "lab_prove", "lab_prove_batch", "lab_verify",
    "lab_witness_load_dev", "lab_commit_inner_dev", "lab_gram_dev", "lab_jl_project_dev", "lab_amortize_z_dev",
    "lab_synth_zq_dev", "lab_synth_pi_dev", "lab_bench_alu_peak",
    "lab_comm_unique_id", "lab_comm_init", "lab_comm_destroy", "lab_transcript_bincode", "lab_crs_cache_configure", "lab_crs_cache_stats",
    "lab_sample_challenge_polys_dev", "lab_generate_witness_dev", "lab_generate_state_dev",
    "lab_comm_allreduce_i64_dev", "lab_comm_allgather_dev", "lab_comm_rank", "lab_comm_shard",
    "lab_pi_pack", "lab_pi_unpack", "lab_pi_pack_dev", "lab_jl_project2", "lab_jl_project2_part", "lab_aggregate_phi2",
    "lab_gram_part", "lab_amortize_z_part", "lab_jl_project2_dev", "lab_jl_project_sharded_dev", "lab_amortize_z_sharded_dev",
    "lab_gram_sharded_dev", "lab_witness_load", "lab_commit_inner_resident", "lab_synth_pi2_dev",
    "lab_transcript_size_in_bytes", "lab_transcript_pack", "lab_transcript_unpack", "lab_fs_init", "lab_fs_absorb", "lab_fs_squeeze",
    "lab_prove_fs", "lab_verify_fs", "lab_rq_add_batch", "lab_rq_sub_batch", "lab_graph_stats",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.abspath(SO_PATH)
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `python __graft_entry__.py` "
                              "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
        _lib = C.CDLL(path)
        _lib.lab_last_error.restype = C.c_char_p
        _lib.lab_last_error.argtypes = [C.c_void_p]
        _lib.lab_stream.restype = C.c_void_p
        _lib.lab_stream.argtypes = [C.c_void_p]
        _lib.lab_kernel_launches.restype = C.c_uint64
        _lib.lab_kernel_launches.argtypes = [C.c_void_p]
        _lib.lab_ctx_destroy.argtypes = [C.c_void_p]
        _lib.lab_ctx_destroy.restype = None
    return _lib
