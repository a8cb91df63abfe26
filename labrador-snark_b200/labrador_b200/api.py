"""Host-side mirror of the reference's public API for the prover path, over the C ABI.

Names follow the reference (constants.rs / structs.rs / proofgen.rs): RuntimeConstants, CRS, State,
Verifier (as the source of challenges), Prover.proof_gen, Transcript.  The reference's toolchain (Rust)
is not present in this image, so this ctypes harness is the caller that can run here; the Rust wrapper a
maintainer would use is in INTEGRATION.md / rust/.  All arithmetic happens in liblabrador_b200.so on the
GPU; numpy is only used to hold buffers.
"""
import ctypes as C

import numpy as np

from . import _lib, synth
from ._lib import Constants, D, JL_ROWS, LabError, Q


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack_pi(pi):
    """int8 JL entries {-1,0,1} -> 2-bit packed words (bit k = +1, bit 16 + k = -1 of the word of 16 entries); same leading
    shape, last axis / 16.  Host-side marshalling (lab_pi_pack)."""
    pi = np.ascontiguousarray(pi, dtype=np.int8)
    if pi.shape[-1] % 16:
        raise ValueError("row length must be a multiple of 16")
    out = np.empty(pi.shape[:-1] + (pi.shape[-1] // 16,), np.uint32)
    rc = _lib.lib().lab_pi_pack(_p(pi), C.c_size_t(pi.size), _p(out))
    if rc != 0:
        raise LabError(rc, "lab_pi_pack: entries must be in {-1,0,1}")
    return out


def unpack_pi(pi2):
    pi2 = np.ascontiguousarray(pi2, dtype=np.uint32)
    out = np.empty(pi2.shape[:-1] + (pi2.shape[-1] * 16,), np.int8)
    rc = _lib.lib().lab_pi_unpack(_p(pi2), C.c_size_t(out.size), _p(out))
    if rc != 0:
        raise LabError(rc, "lab_pi_unpack failed")
    return out


def _chal(ch, keep):
    """CChallenges from a challenges dict; uses ch["pi2"] (packed) when present, else ch["pi"] (int8)."""
    pi = pi2 = None
    if ch.get("pi2") is not None:
        pi2 = np.ascontiguousarray(ch["pi2"], dtype=np.uint32)
        if pi2.ndim == 3:
            pi2 = pi2[None]
        n_att = pi2.shape[0]
    if ch.get("pi") is not None and pi2 is None:
        pi = np.ascontiguousarray(ch["pi"], dtype=np.int8)
        if pi.ndim == 3:
            pi = pi[None]
        n_att = pi.shape[0]
    omega, alpha, beta, cc = _u32(ch["omega"]), _u32(ch["alpha"]), _u32(ch["beta"]), _u32(ch["c"])
    keep.extend([pi, pi2, omega, alpha, beta, cc])
    return _lib.CChallenges(_p(pi) if pi is not None else None, n_att, int(ch["psi"]), _p(omega), _p(alpha), _p(beta), _p(cc),
                            _p(pi2) if pi2 is not None else None)


def _seed_buf(seed):
    s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    if s.size != 32:
        raise ValueError("CRS seed must be 32 bytes")
    return s


class RuntimeConstants:
    """RuntimeConstants::new(N, R) (constants.rs:234-264)."""

    @staticmethod
    def new(N, R, allow_degenerate=False):
        c = Constants()
        rc = _lib.lib().lab_runtime_constants(C.c_uint64(N), C.c_uint64(R), C.byref(c))
        if rc != 0 and not allow_degenerate:
            raise LabError(rc, f"RuntimeConstants::new({N},{R}) is degenerate (B={c.B}, T_1={c.T_1}, T_2={c.T_2})")
        return c


class Context:
    """One CUDA device + stream + scratch arena (lab_ctx)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        L = _lib.lib()
        rc = L.lab_ctx_create(C.c_int(device), C.byref(self._h))
        if rc != 0:
            raise LabError(rc, L.lab_last_error(None).decode())
        self.L = L
        self.device = device

    def close(self):
        if self._h:
            self.L.lab_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LabError(rc, self.L.lab_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return int(self.L.lab_kernel_launches(self._h))

    @property
    def stream(self):
        return self.L.lab_stream(self._h)

    def sync(self):
        self._ck(self.L.lab_sync(self._h))

    def timer_start(self):
        self._ck(self.L.lab_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double(0)
        self._ck(self.L.lab_timer_stop(self._h, C.byref(ms)))
        return ms.value

    # ---- multi-GPU: NCCL communicator inside the library (lab_comm_*) ----
    @staticmethod
    def comm_unique_id():
        """128 bytes that rank 0 generates and the host distributes to all ranks (any transport)."""
        buf = (C.c_uint8 * 128)()
        L = _lib.lib()
        rc = L.lab_comm_unique_id(buf)
        if rc != 0:
            raise LabError(rc, L.lab_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, uid, rank, world):
        """Collective.  Afterwards proof_gen / verify called by all ranks with the same arguments shard the
        CRS-regenerating stages by rows and all-gather them over NVLink; every rank gets the same transcript."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid))
        self._ck(self.L.lab_comm_init(self._h, buf, C.c_int(rank), C.c_int(world)))
        self.rank, self.world = rank, world

    def comm_destroy(self):
        self._ck(self.L.lab_comm_destroy(self._h))

    # ---- device-pointer entry points (inputs resident in HBM) ----
    def ntt_fwd_batch_dev(self, din, dout, n):
        self._ck(self.L.lab_ntt_fwd_batch_dev(self._h, C.c_void_p(din), C.c_void_p(dout), C.c_size_t(n)))

    def ntt_inv_batch_dev(self, din, dout, n):
        self._ck(self.L.lab_ntt_inv_batch_dev(self._h, C.c_void_p(din), C.c_void_p(dout), C.c_size_t(n)))

    def polymul_batch_dev(self, da, db, dc, n):
        self._ck(self.L.lab_polymul_batch_dev(self._h, C.c_void_p(da), C.c_void_p(db), C.c_void_p(dc), C.c_size_t(n)))

    def crs_expand_dev(self, seed, start, n_polys, dout):
        self._ck(self.L.lab_crs_expand_dev(self._h, _p(_seed_buf(seed)), C.c_uint64(start & (2**64 - 1)), C.c_uint64(start >> 64),
                                           C.c_size_t(n_polys), C.c_void_p(dout)))

    def witness_load_dev(self, c, dS):
        self._ck(self.L.lab_witness_load_dev(self._h, C.byref(c), C.c_void_p(dS)))

    def commit_inner_dev(self, seed, row0, nrows, dT):
        self._ck(self.L.lab_commit_inner_dev(self._h, _p(_seed_buf(seed)), C.c_uint64(row0), C.c_uint64(nrows), C.c_void_p(dT)))

    def gram_dev(self, i0, ni, dG):
        self._ck(self.L.lab_gram_dev(self._h, C.c_uint64(i0), C.c_uint64(ni), C.c_void_p(dG)))

    def synth_zq_dev(self, seed, stream, start, n, dout):
        self._ck(self.L.lab_synth_zq_dev(self._h, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(start), C.c_size_t(n), C.c_void_p(dout)))

    def synth_pi_dev(self, seed, attempt, first_entry, total, dout):
        self._ck(self.L.lab_synth_pi_dev(self._h, C.c_uint64(seed), C.c_uint64(attempt), C.c_uint64(first_entry), C.c_size_t(total), C.c_void_p(dout)))

    # ---- CRS cache (lab_crs_cache_*) ----
    def crs_cache_configure(self, max_bytes):
        """Keep the transformed CRS polynomials of the outer commitments in HBM (up to max_bytes) so that verify after
        prove, and further proofs under the same CRS, skip ChaCha20.  0 disables and frees.  Bit-identical results."""
        self._ck(self.L.lab_crs_cache_configure(self._h, C.c_size_t(max_bytes)))

    def crs_cache_stats(self):
        used, hits, misses = C.c_size_t(0), C.c_uint64(0), C.c_uint64(0)
        self._ck(self.L.lab_crs_cache_stats(self._h, C.byref(used), C.byref(hits), C.byref(misses)))
        return {"bytes": used.value, "hits": hits.value, "misses": misses.value}

    def graph_stats(self):
        g, r, f = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        self._ck(self.L.lab_graph_stats(self._h, C.byref(g), C.byref(r), C.byref(f)))
        return {"graphs": g.value, "replays": r.value, "failed": bool(f.value)}

    # ---- device-side generation (f2 challenges, f4 witness / statement) ----
    def sample_challenge_polys(self, seed, first_idx, count):
        """Verifier::fetch_challenge on the device -> ([count][64] canonical, candidates tried per polynomial)."""
        dc, dn = self.malloc(count * D * 4), self.malloc(count * 4)
        try:
            self._ck(self.L.lab_sample_challenge_polys_dev(self._h, C.c_uint64(seed), C.c_uint32(first_idx), C.c_uint32(count), C.c_void_p(dc), C.c_void_p(dn)))
            out, tries = np.empty((count, D), np.uint32), np.empty(count, np.uint32)
            self.d2h(out, dc); self.d2h(tries, dn); self.sync()
        finally:
            self.free(dc); self.free(dn)
        return out, tries

    def generate_witness_dev(self, c, seed, dS):
        info = (C.c_uint64 * 2)()
        self._ck(self.L.lab_generate_witness_dev(self._h, C.byref(c), C.c_uint64(seed), C.c_void_p(dS), info))
        return int(info[0]), int(info[1])

    def generate_witness(self, c, seed):
        """generate_witness (proofgen.rs:460-518), seeded, computed on the device -> [R][N][64]."""
        n = c.R * c.N * D
        dS = self.malloc(n * 4)
        try:
            norm, draws = self.generate_witness_dev(c, seed, dS)
            S = np.empty((c.R, c.N, D), np.uint32)
            self.d2h(S, dS); self.sync()
        finally:
            self.free(dS)
        return S, norm, draws

    def generate_state(self, c, seed, S):
        """State::gen_f (structs.rs:289-350), seeded, computed on the device -> (phi, a, b)."""
        S = _u32(S)
        dS, dphi, da, db = self.malloc(S.nbytes), self.malloc(S.nbytes), self.malloc(c.R * c.R * D * 4), self.malloc(D * 4)
        try:
            self.h2d(dS, S)
            self._ck(self.L.lab_generate_state_dev(self._h, C.byref(c), C.c_uint64(seed), C.c_void_p(dS), C.c_void_p(dphi), C.c_void_p(da), C.c_void_p(db)))
            phi, a, b = np.empty((c.R, c.N, D), np.uint32), np.empty((c.R, c.R, D), np.uint32), np.empty(D, np.uint32)
            self.d2h(phi, dphi); self.d2h(a, da); self.d2h(b, db); self.sync()
        finally:
            for d in (dS, dphi, da, db):
                self.free(d)
        return phi, a, b

    def alu_peak(self):
        v = C.c_double(0)
        self._ck(self.L.lab_bench_alu_peak(self._h, C.byref(v)))
        return v.value

    def jl_project_dev(self, dpi, i0, ni, dp):
        self._ck(self.L.lab_jl_project_dev(self._h, C.c_void_p(dpi), C.c_uint64(i0), C.c_uint64(ni), C.c_void_p(dp)))

    def jl_project2_dev(self, dpi2, i0, ni, dp):
        self._ck(self.L.lab_jl_project2_dev(self._h, C.c_void_p(dpi2), C.c_uint64(i0), C.c_uint64(ni), C.c_void_p(dp)))

    def pi_pack_dev(self, dpi, n_entries, dpi2):
        self._ck(self.L.lab_pi_pack_dev(self._h, C.c_void_p(dpi), C.c_size_t(n_entries), C.c_void_p(dpi2)))

    def synth_pi2_dev(self, seed, attempt, first_entry, total, dout):
        self._ck(self.L.lab_synth_pi2_dev(self._h, C.c_uint64(seed), C.c_uint64(attempt), C.c_uint64(first_entry), C.c_size_t(total), C.c_void_p(dout)))

    # ---- stage calls sharded over the library's communicator (all ranks call them) ----
    def comm_shard(self, total):
        x0, nx = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.L.lab_comm_shard(self._h, C.c_uint64(total), C.byref(x0), C.byref(nx)))
        return x0.value, nx.value

    def comm_allreduce_i64_dev(self, dbuf, n):
        self._ck(self.L.lab_comm_allreduce_i64_dev(self._h, C.c_void_p(dbuf), C.c_size_t(n)))

    def comm_allgather_dev(self, dbuf, bytes_per_rank):
        self._ck(self.L.lab_comm_allgather_dev(self._h, C.c_void_p(dbuf), C.c_size_t(bytes_per_rank)))

    def jl_project_sharded_dev(self, dpi2_part, dp):
        self._ck(self.L.lab_jl_project_sharded_dev(self._h, C.c_void_p(dpi2_part), C.c_void_p(dp)))

    def amortize_z_sharded_dev(self, dch, dz):
        self._ck(self.L.lab_amortize_z_sharded_dev(self._h, C.c_void_p(dch), C.c_void_p(dz)))

    def gram_sharded_dev(self, dG):
        self._ck(self.L.lab_gram_sharded_dev(self._h, C.c_void_p(dG)))

    def witness_load(self, c, S):
        S = _u32(S)
        self._ck(self.L.lab_witness_load(self._h, C.byref(c), _p(S)))
        self.sync()

    def amortize_z_dev(self, dch, i0, ni, dz):
        self._ck(self.L.lab_amortize_z_dev(self._h, C.c_void_p(dch), C.c_uint64(i0), C.c_uint64(ni), C.c_void_p(dz)))

    def norm_sq_dev(self, din, n):
        out = C.c_uint64(0)
        self._ck(self.L.lab_norm_sq_dev(self._h, C.c_void_p(din), C.c_size_t(n), C.byref(out)))
        return out.value

    # ---- raw device memory (device-resident API) ----
    def malloc(self, nbytes):
        p = C.c_void_p()
        self._ck(self.L.lab_malloc(self._h, C.c_size_t(nbytes), C.byref(p)))
        return p.value

    def free(self, dptr):
        self._ck(self.L.lab_free(self._h, C.c_void_p(dptr)))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(self.L.lab_memcpy_h2d(self._h, C.c_void_p(dptr), _p(arr), C.c_size_t(arr.nbytes)))
        return arr

    def d2h(self, arr, dptr):
        self._ck(self.L.lab_memcpy_d2h(self._h, _p(arr), C.c_void_p(dptr), C.c_size_t(arr.nbytes)))

    # ---- ring primitives ----
    def ntt_fwd_batch(self, polys):
        a = _u32(polys).reshape(-1, D)
        out = np.empty_like(a)
        self._ck(self.L.lab_ntt_fwd_batch(self._h, _p(a), _p(out), C.c_size_t(a.shape[0])))
        return out

    def ntt_inv_batch(self, slots):
        a = _u32(slots).reshape(-1, D)
        out = np.empty_like(a)
        self._ck(self.L.lab_ntt_inv_batch(self._h, _p(a), _p(out), C.c_size_t(a.shape[0])))
        return out

    def polymul_batch(self, a, b):
        a, b = _u32(a).reshape(-1, D), _u32(b).reshape(-1, D)
        if a.shape != b.shape:
            raise LabError(3, "polymul_batch: shape mismatch")
        out = np.empty_like(a)
        self._ck(self.L.lab_polymul_batch(self._h, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0])))
        return out

    def rq_add_batch(self, a, b, sub=False):
        """&Rq + &Rq / &Rq - &Rq (algebraic.rs:441-515), batched."""
        a, b = _u32(a).reshape(-1, D), _u32(b).reshape(-1, D)
        if a.shape != b.shape:
            raise LabError(3, "rq_add_batch: shape mismatch")
        out = np.empty_like(a)
        f = self.L.lab_rq_sub_batch if sub else self.L.lab_rq_add_batch
        self._ck(f(self._h, _p(a), _p(b), _p(out), C.c_size_t(a.shape[0])))
        return out

    def inner_product_batch(self, v1, v2):
        """polynomial_vec_inner_product (util.rs:496-509) for a batch: v1, v2 [B][len][64] -> [B][64]."""
        v1, v2 = _u32(v1), _u32(v2)
        if v1.shape != v2.shape:   # util.rs:497-502 assert
            raise LabError(3, f"inner product not defined on vectors of unequal length. v1 length: {v1.shape}, v2 length: {v2.shape}")
        B, ln = v1.shape[0], v1.shape[1]
        out = np.empty((B, D), np.uint32)
        self._ck(self.L.lab_inner_product_batch(self._h, _p(v1), _p(v2), C.c_size_t(B), C.c_size_t(ln), _p(out)))
        return out

    def decompose(self, polys, base, exp):
        a = _u32(polys).reshape(-1, D)
        out = np.empty((exp, a.shape[0], D), np.uint32)
        self._ck(self.L.lab_decompose(self._h, _p(a), C.c_size_t(a.shape[0]), C.c_int64(base), C.c_int64(exp), _p(out)))
        return out

    def norm_sq(self, x):
        a = _u32(x).reshape(-1)
        out = C.c_uint64(0)
        self._ck(self.L.lab_norm_sq(self._h, _p(a), C.c_size_t(a.size), C.byref(out)))
        return out.value

    def sigma_inv(self, polys):
        a = _u32(polys).reshape(-1, D)
        out = np.empty_like(a)
        self._ck(self.L.lab_sigma_inv(self._h, _p(a), C.c_size_t(a.shape[0]), _p(out)))
        return out

    # ---- CRS ----
    def crs_expand(self, seed, start, n_polys):
        s = _seed_buf(seed)
        out = np.empty((n_polys, D), np.uint32)
        self._ck(self.L.lab_crs_expand(self._h, _p(s), C.c_uint64(start & (2**64 - 1)), C.c_uint64(start >> 64), C.c_size_t(n_polys), _p(out)))
        return out

    # ---- stages ----
    def commit_inner(self, c, seed, S, row0=0, nrows=None):
        S = _u32(S)
        nrows = c.KAPPA - row0 if nrows is None else nrows
        T = np.empty((c.R, nrows, D), np.uint32)
        self._ck(self.L.lab_commit_inner(self._h, C.byref(c), _p(_seed_buf(seed)), _p(S), C.c_uint64(row0), C.c_uint64(nrows), _p(T)))
        return T

    def gram(self, c, S):
        S = _u32(S)
        G = np.empty((c.R, c.R, D), np.uint32)
        self._ck(self.L.lab_gram(self._h, C.byref(c), _p(S), _p(G)))
        return G

    def jl_project(self, c, S, pi):
        S = _u32(S)
        pi = np.ascontiguousarray(pi, dtype=np.int8)
        p = np.empty(JL_ROWS, np.int64)
        acc = C.c_int(0)
        self._ck(self.L.lab_jl_project(self._h, C.byref(c), _p(S), _p(pi), _p(p), C.byref(acc)))
        return p, bool(acc.value)

    def jl_project2(self, c, S, pi2):
        S = _u32(S)
        pi2 = np.ascontiguousarray(pi2, dtype=np.uint32)
        p = np.empty(JL_ROWS, np.int64)
        acc = C.c_int(0)
        self._ck(self.L.lab_jl_project2(self._h, C.byref(c), _p(S), _p(pi2), _p(p), C.byref(acc)))
        return p, bool(acc.value)

    def jl_project2_part(self, c, S, pi2_part, i0, ni):
        S = _u32(S)
        pi2_part = np.ascontiguousarray(pi2_part, dtype=np.uint32)
        p = np.empty(JL_ROWS, np.int64)
        self._ck(self.L.lab_jl_project2_part(self._h, C.byref(c), _p(S), _p(pi2_part), C.c_uint64(i0), C.c_uint64(ni), _p(p)))
        return p

    def aggregate_phi2(self, c, phi, pi2, psi, omega):
        phi, omega = _u32(phi), _u32(omega)
        pi2 = np.ascontiguousarray(pi2, dtype=np.uint32)
        out = np.empty((c.R, c.N, D), np.uint32)
        self._ck(self.L.lab_aggregate_phi2(self._h, C.byref(c), _p(phi), _p(pi2), C.c_uint32(psi), _p(omega), _p(out)))
        return out

    def gram_part(self, c, S, i0, ni):
        S = _u32(S)
        G = np.empty((ni, c.R, D), np.uint32)
        self._ck(self.L.lab_gram_part(self._h, C.byref(c), _p(S), C.c_uint64(i0), C.c_uint64(ni), _p(G)))
        return G

    def amortize_z_part(self, c, S, ch, i0, ni):
        S, ch = _u32(S), _u32(ch)
        z = np.empty((c.N, D), np.uint32)
        self._ck(self.L.lab_amortize_z_part(self._h, C.byref(c), _p(S), _p(ch), C.c_uint64(i0), C.c_uint64(ni), _p(z)))
        return z

    def jl_project_part(self, c, S, pi_part, i0, ni):
        S = _u32(S)
        pi_part = np.ascontiguousarray(pi_part, dtype=np.int8)
        p = np.empty(JL_ROWS, np.int64)
        self._ck(self.L.lab_jl_project_part(self._h, C.byref(c), _p(S), _p(pi_part), C.c_uint64(i0), C.c_uint64(ni), _p(p)))
        return p

    def commit_outer_u1(self, c, seed, T, G):
        T, G = _u32(T), _u32(G)
        u1 = np.empty((c.KAPPA_1, D), np.uint32)
        self._ck(self.L.lab_commit_outer_u1(self._h, C.byref(c), _p(_seed_buf(seed)), _p(T), _p(G), _p(u1)))
        return u1

    def commit_outer_u2(self, c, seed, H):
        H = _u32(H)
        u2 = np.empty((c.KAPPA_2, D), np.uint32)
        self._ck(self.L.lab_commit_outer_u2(self._h, C.byref(c), _p(_seed_buf(seed)), _p(H), _p(u2)))
        return u2

    def aggregate_phi(self, c, phi, pi, psi, omega):
        phi, omega = _u32(phi), _u32(omega)
        pi = np.ascontiguousarray(pi, dtype=np.int8)
        out = np.empty((c.R, c.N, D), np.uint32)
        self._ck(self.L.lab_aggregate_phi(self._h, C.byref(c), _p(phi), _p(pi), C.c_uint32(psi), _p(omega), _p(out)))
        return out

    def h_gram(self, c, phi_final, S):
        phi_final, S = _u32(phi_final), _u32(S)
        H = np.empty((c.R, c.R, D), np.uint32)
        self._ck(self.L.lab_h_gram(self._h, C.byref(c), _p(phi_final), _p(S), _p(H)))
        return H

    def amortize_z(self, c, S, ch):
        S, ch = _u32(S), _u32(ch)
        z = np.empty((c.N, D), np.uint32)
        self._ck(self.L.lab_amortize_z(self._h, C.byref(c), _p(S), _p(ch), _p(z)))
        return z


    # ---- Verifier::verify on the GPU and batched proving ----
    def verify(self, c, seed, phi, a, b, ch, tr):
        """tr: dict with u_1, projection_int, projection, b_prime_prime, u_2, z, t, g, h, jl_attempt (oracle layout).
        Returns (accepted, failed_check, norm_sum) like the reference's Verifier::verify (verification.rs:25-438)."""
        phi, a, b = _u32(phi), _u32(a), _u32(b)
        alive = []
        cch = _chal(ch, alive)
        keep = {k: _u32(tr[k]) for k in ("u_1", "projection", "b_prime_prime", "u_2", "z", "t", "g", "h")}
        pint = np.ascontiguousarray(tr["projection_int"], dtype=np.int64)
        cst = _lib.CState(_p(phi), _p(a), _p(b))
        ctr = _lib.CTranscript(_p(keep["u_1"]), int(tr.get("jl_attempt", 0)), _p(pint), _p(keep["projection"]), _p(keep["b_prime_prime"]),
                               _p(keep["u_2"]), _p(keep["z"]), _p(keep["t"]), _p(keep["g"]), _p(keep["h"]), None, 0)
        acc, fc, ns = C.c_int(0), C.c_int(0), C.c_uint64(0)
        self._ck(self.L.lab_verify(self._h, C.byref(c), _p(_seed_buf(seed)), C.byref(cst), C.byref(cch), C.byref(ctr),
                                   C.byref(acc), C.byref(fc), C.byref(ns)))
        return bool(acc.value), fc.value, ns.value

    # ---- Fiat-Shamir: the verifier's randomness derived from the transcript (lab_prove_fs / lab_verify_fs) ----
    def prove_fs(self, c, seed, S, phi, a, b):
        """Returns (transcript dict, challenges dict as derived: accepted attempt packed in pi2[1], jl_attempt = tries before it)."""
        S, phi, a, b = _u32(S), _u32(phi), _u32(a), _u32(b)
        o = _alloc_out(c)
        ctr = _ctr(o)
        bufs, cb = _alloc_chbuf(c)
        cst = _lib.CState(_p(phi), _p(a), _p(b))
        rc = self.L.lab_prove_fs(self._h, C.byref(c), _p(_seed_buf(seed)), _p(S), C.byref(cst), C.byref(ctr), C.byref(cb))
        if rc == 1:
            raise LabError(rc, "failed JL...")
        self._ck(rc)
        o["jl_attempt"] = ctr.jl_attempt
        o["norm_sum"] = int(ctr.norm_sum)
        ch = {"pi": None, "pi2": bufs["pi2"][None], "psi": int(cb.psi), "omega": bufs["omega"], "alpha": bufs["alpha"], "beta": bufs["beta"], "c": bufs["c"]}
        return o, ch

    def verify_fs(self, c, seed, phi, a, b, tr):
        phi, a, b = _u32(phi), _u32(a), _u32(b)
        keep = {k: _u32(tr[k]) for k in ("u_1", "projection", "b_prime_prime", "u_2", "z", "t", "g", "h")}
        pint = np.ascontiguousarray(tr["projection_int"], dtype=np.int64)
        cst = _lib.CState(_p(phi), _p(a), _p(b))
        ctr = _lib.CTranscript(_p(keep["u_1"]), int(tr.get("jl_attempt", 0)), _p(pint), _p(keep["projection"]), _p(keep["b_prime_prime"]),
                               _p(keep["u_2"]), _p(keep["z"]), _p(keep["t"]), _p(keep["g"]), _p(keep["h"]), None, 0)
        acc, fc, ns = C.c_int(0), C.c_int(0), C.c_uint64(0)
        self._ck(self.L.lab_verify_fs(self._h, C.byref(c), _p(_seed_buf(seed)), C.byref(cst), C.byref(ctr), C.byref(acc), C.byref(fc), C.byref(ns)))
        return bool(acc.value), fc.value, ns.value

    def prove_batch(self, c, seeds, shared_crs, S, phi, a, b, challenges):
        """lab_prove_batch: S [B][R][N][64], phi [B][R][N][64], a [B][R][R][64], b [B][64]; challenges: list of dicts.
        Returns a list of oracle-layout transcript dicts."""
        S, phi, a, b = _u32(S), _u32(phi), _u32(a), _u32(b)
        B = S.shape[0]
        seedbuf = np.frombuffer(b"".join(bytes(s) for s in seeds), dtype=np.uint8).copy()
        sts = (_lib.CState * B)()
        chs = (_lib.CChallenges * B)()
        trs = (_lib.CTranscript * B)()
        keep, outs = [], []
        for i in range(B):
            ch = challenges[i]
            arrs = []
            chs[i] = _chal(ch, arrs)
            keep.append(arrs)
            sts[i] = _lib.CState(_p(phi[i]), _p(a[i]), _p(b[i]))
            o = {"u_1": np.zeros((c.KAPPA_1, D), np.uint32), "projection_int": np.zeros(JL_ROWS, np.int64), "projection": np.zeros(JL_ROWS, np.uint32),
                 "b_prime_prime": np.zeros(D, np.uint32), "u_2": np.zeros((c.KAPPA_2, D), np.uint32), "z": np.zeros((c.N, D), np.uint32),
                 "t": np.zeros((c.R, c.KAPPA, D), np.uint32), "g": np.zeros((c.R, c.R, D), np.uint32), "h": np.zeros((c.R, c.R, D), np.uint32),
                 "phi_final": np.zeros((c.R, c.N, D), np.uint32)}
            outs.append(o)
            trs[i] = _lib.CTranscript(_p(o["u_1"]), 0, _p(o["projection_int"]), _p(o["projection"]), _p(o["b_prime_prime"]), _p(o["u_2"]),
                                      _p(o["z"]), _p(o["t"]), _p(o["g"]), _p(o["h"]), _p(o["phi_final"]), 0)
        import time as _time
        t0 = _time.perf_counter()
        rc = self.L.lab_prove_batch(self._h, C.byref(c), C.c_size_t(B), _p(seedbuf), C.c_int(int(shared_crs)), _p(S), sts, chs, trs)
        self.last_batch_seconds = _time.perf_counter() - t0        # the C call alone (host buffers in, transcripts out), without the ctypes marshalling above
        self._ck(rc)
        for i in range(B):
            outs[i]["jl_attempt"] = trs[i].jl_attempt
            outs[i]["norm_sum"] = int(trs[i].norm_sum)
        return outs


def transcript_bincode(c, tr, ch, jl_attempt=None):
    """tr: oracle-layout transcript dict; ch: challenges dict (pi [attempts][R][256][N*64]).  Host-only: needs no GPU."""
    L = _lib.lib()
    alive = []
    cch = _chal(ch, alive)
    keep = {k: _u32(tr[k]) for k in ("u_1", "projection", "b_prime_prime", "u_2", "z", "t", "g", "h")}
    att = int(tr.get("jl_attempt", 0)) if jl_attempt is None else jl_attempt
    ctr = _lib.CTranscript(_p(keep["u_1"]), att, None, _p(keep["projection"]), _p(keep["b_prime_prime"]),
                           _p(keep["u_2"]), _p(keep["z"]), _p(keep["t"]), _p(keep["g"]), _p(keep["h"]), None, 0)
    size = C.c_size_t(0)
    rc = L.lab_transcript_bincode(C.byref(c), C.byref(ctr), C.byref(cch), None, C.c_size_t(0), C.byref(size))
    if rc != 0:
        raise LabError(rc, "lab_transcript_bincode: bad arguments")
    buf = np.empty(size.value, np.uint8)
    rc = L.lab_transcript_bincode(C.byref(c), C.byref(ctr), C.byref(cch), _p(buf), C.c_size_t(buf.size), C.byref(size))
    if rc != 0:
        raise LabError(rc, "lab_transcript_bincode failed")
    return buf.tobytes()


def transcript_size_in_bytes(c, tr, ch):
    """Transcript::size_in_bytes (structs.rs:211-221): (gzip-compressed size, bincode size) in bytes."""
    alive = []
    cch = _chal(ch, alive)
    keep = {k: _u32(tr[k]) for k in ("u_1", "projection", "b_prime_prime", "u_2", "z", "t", "g", "h")}
    ctr = _lib.CTranscript(_p(keep["u_1"]), int(tr.get("jl_attempt", 0)), None, _p(keep["projection"]), _p(keep["b_prime_prime"]),
                           _p(keep["u_2"]), _p(keep["z"]), _p(keep["t"]), _p(keep["g"]), _p(keep["h"]), None, 0)
    gz, raw = C.c_size_t(0), C.c_size_t(0)
    rc = _lib.lib().lab_transcript_size_in_bytes(C.byref(c), C.byref(ctr), C.byref(cch), C.byref(gz), C.byref(raw))
    if rc != 0:
        raise LabError(rc, "lab_transcript_size_in_bytes failed")
    return gz.value, raw.value


def transcript_pack(c, tr, ch):
    """Compact wire format (13-bit coefficients, 2-bit JL entries): bytes."""
    L = _lib.lib()
    alive = []
    cch = _chal(ch, alive)
    keep = {k: _u32(tr[k]) for k in ("u_1", "projection", "b_prime_prime", "u_2", "z", "t", "g", "h")}
    ctr = _lib.CTranscript(_p(keep["u_1"]), int(tr.get("jl_attempt", 0)), None, _p(keep["projection"]), _p(keep["b_prime_prime"]),
                           _p(keep["u_2"]), _p(keep["z"]), _p(keep["t"]), _p(keep["g"]), _p(keep["h"]), None, 0)
    size = C.c_size_t(0)
    rc = L.lab_transcript_pack(C.byref(c), C.byref(ctr), C.byref(cch), None, C.c_size_t(0), C.byref(size))
    if rc != 0:
        raise LabError(rc, "lab_transcript_pack: bad arguments (g and h must be symmetric)")
    buf = np.empty(size.value, np.uint8)
    rc = L.lab_transcript_pack(C.byref(c), C.byref(ctr), C.byref(cch), _p(buf), C.c_size_t(buf.size), C.byref(size))
    if rc != 0:
        raise LabError(rc, "lab_transcript_pack failed")
    return buf.tobytes()


def _alloc_out(c):
    return {"u_1": np.zeros((c.KAPPA_1, D), np.uint32), "projection_int": np.zeros(JL_ROWS, np.int64), "projection": np.zeros(JL_ROWS, np.uint32),
            "b_prime_prime": np.zeros(D, np.uint32), "u_2": np.zeros((c.KAPPA_2, D), np.uint32), "z": np.zeros((c.N, D), np.uint32),
            "t": np.zeros((c.R, c.KAPPA, D), np.uint32), "g": np.zeros((c.R, c.R, D), np.uint32), "h": np.zeros((c.R, c.R, D), np.uint32),
            "phi_final": np.zeros((c.R, c.N, D), np.uint32)}


def _ctr(o, jl_attempt=0):
    return _lib.CTranscript(_p(o["u_1"]), jl_attempt, _p(o["projection_int"]), _p(o["projection"]), _p(o["b_prime_prime"]), _p(o["u_2"]),
                            _p(o["z"]), _p(o["t"]), _p(o["g"]), _p(o["h"]), _p(o["phi_final"]), 0)


def _alloc_chbuf(c):
    b = {"pi2": np.zeros((c.R, JL_ROWS, c.N * 4), np.uint32), "omega": np.zeros(JL_ROWS, np.uint32), "alpha": np.zeros(D, np.uint32),
         "beta": np.zeros(D, np.uint32), "c": np.zeros((c.R, D), np.uint32)}
    return b, _lib.CChallengesBuf(_p(b["pi2"]), 0, _p(b["omega"]), _p(b["alpha"]), _p(b["beta"]), _p(b["c"]))


def transcript_unpack(c, blob):
    """Inverse of transcript_pack -> (transcript dict, challenges dict with the accepted attempt as pi2[1][R][256][N*4])."""
    data = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
    o = _alloc_out(c)
    ctr = _ctr(o)
    b, cb = _alloc_chbuf(c)
    rc = _lib.lib().lab_transcript_unpack(C.byref(c), _p(data), C.c_size_t(data.size), C.byref(ctr), C.byref(cb))
    if rc != 0:
        raise LabError(rc, "lab_transcript_unpack: malformed input")
    o["jl_attempt"] = ctr.jl_attempt
    ch = {"pi": None, "pi2": b["pi2"][None], "psi": int(cb.psi), "omega": b["omega"], "alpha": b["alpha"], "beta": b["beta"], "c": b["c"]}
    return o, ch


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class CRS:
    """CRS (structs.rs:27-190).  `from_seed` is the addition parity needs (the reference's CRS::new draws
    the seed from thread_rng and keeps it private, structs.rs:173-189)."""

    def __init__(self, constants, base_seed, ctx=None):
        self.constants = constants
        self.base_seed = bytes(base_seed)
        self.ctx = ctx or default_context()

    @classmethod
    def new(cls, constants, ctx=None):
        import os
        return cls(constants, os.urandom(32), ctx)

    @classmethod
    def from_seed(cls, constants, seed, ctx=None):
        return cls(constants, seed, ctx)

    def _fetch(self, which, i=0, j=0, k=0, row=0):
        c = self.constants
        n = c.N if which == "A" else (c.KAPPA if which == "B" else c.KAPPA_2)
        out = np.empty((n, D), np.uint32)
        self.ctx._ck(self.ctx.L.lab_crs_fetch(self.ctx._h, C.byref(c), _p(_seed_buf(self.base_seed)), C.c_int(ord(which)),
                                              C.c_uint64(i), C.c_uint64(j), C.c_uint64(k), C.c_uint64(row), _p(out)))
        return out

    def fetch_A_row(self, row):
        return self._fetch("A", row=row)

    def fetch_B_ik_row(self, i, k, row):
        return self._fetch("B", i=i, k=k, row=row)

    def fetch_C_ijk(self, i, j, k):
        return self._fetch("C", i=i, j=j, k=k)

    def fetch_D_ijk(self, i, j, k):
        return self._fetch("D", i=i, j=j, k=k)


class State:
    """State (structs.rs:269-388) with K = L = 1: phi [R][N][64], a [R][R][64] symmetric, b [64]."""

    def __init__(self, phi, a, b):
        self.phi_k = [_u32(phi)]
        self.a_k = [_u32(a)]
        self.b_k = [_u32(b)]
        self.phi_prime_k, self.a_prime_k = self.phi_k, self.a_k          # structs.rs:367-372
        self.b_prime_k = [int(self.b_k[0][0])]                           # b.eval(0), structs.rs:373

    @classmethod
    def new(cls, witness, constants, seed=synth.SEED, ctx=None):
        """gen_f (structs.rs:289-350): random symmetric a, random phi, b = sum a_ij <s_i,s_j> + sum <phi_i,s_i>."""
        ctx = ctx or default_context()
        c = constants
        S = _u32(witness)
        phi, a = synth.generate_statement_inputs(c.N, c.R, seed)
        G = ctx.gram(c, S)
        ag = ctx.polymul_batch(a.reshape(-1, D), G.reshape(-1, D)).astype(np.uint64).sum(axis=0)
        ps = ctx.inner_product_batch(phi, S).astype(np.uint64).sum(axis=0)
        b = ((ag + ps) % Q).astype(np.uint32)
        return cls(phi, a, b)


class Verifier:
    """The part of Verifier the prover talks to (verification.rs:441-513,553-579): a replayable challenge
    source.  Challenges may be given explicitly (dict as synth.sample_challenges returns) or drawn from the
    seeded PRG."""

    def __init__(self, b_prime_k, constants, challenges=None, seed=synth.SEED, n_attempts=6):
        self.b_prime = list(b_prime_k)
        self.constants = constants
        c = constants
        self.challenges = challenges if challenges is not None else synth.sample_challenges(c.N, c.R, seed, n_attempts)

    @classmethod
    def new(cls, b_prime_k, constants, **kw):
        return cls(b_prime_k, constants, **kw)

    def sample_jl_projection(self, attempt=0):
        return self.challenges["pi"][attempt]

    def generate_psi(self):
        return [self.challenges["psi"]]

    def generate_omega(self):
        return self.challenges["omega"]

    def fetch_alpha(self):
        return [self.challenges["alpha"]]

    def fetch_beta(self):
        return [self.challenges["beta"]]

    def fetch_challenge(self, i):
        return self.challenges["c"][i]


class Transcript:
    """Transcript (structs.rs:192-209), dense arrays.  pi_i_all is the accepted JL attempt lifted to Z_q."""

    FIELDS = ("u_1", "projection", "psi", "omega", "b_prime_prime", "alpha", "beta", "u_2", "c", "z", "t_i_all", "g_mat", "h_mat")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def pi_i_all(self):
        pi = self.pi_accepted.astype(np.int32)
        return np.where(pi < 0, pi + Q, pi).astype(np.uint32)

    def to_bincode(self, constants):
        """bincode::serialize(&Transcript) of the reference (structs.rs:192-221), byte for byte (lab_transcript_bincode)."""
        return transcript_bincode(constants, self.as_oracle_dict(),
                                  {"pi": self.pi_accepted[None], "psi": self.psi[0][0], "omega": self.omega[0], "alpha": self.alpha[0],
                                   "beta": self.beta[0], "c": self.c}, jl_attempt=0)

    def as_oracle_dict(self):
        return {"u_1": self.u_1, "projection_int": self.projection_int, "projection": self.projection,
                "b_prime_prime": self.b_prime_prime[0], "u_2": self.u_2, "z": self.z, "t": self.t_i_all,
                "g": self.g_mat, "h": self.h_mat, "phi_final": self.phi_final, "jl_attempt": self.jl_attempt}


class Prover:
    """Prover (proofgen.rs:14-28)."""

    def __init__(self, witness, verifier, constants, ctx=None):
        self.witness = _u32(witness)
        self.verifier = verifier
        self.constants = constants
        self.ctx = ctx or default_context()

    @classmethod
    def new(cls, witness, verifier, constants, ctx=None):
        return cls(witness, verifier, constants, ctx)

    def jl_project(self, attempt=0):
        """Prover::jl_project (proofgen.rs:429-457): (projection ints, Pi lifted)."""
        pi = self.verifier.sample_jl_projection(attempt)
        p, _ = self.ctx.jl_project(self.constants, self.witness, pi)
        lifted = np.where(pi < 0, pi.astype(np.int32) + Q, pi).astype(np.uint32)
        return p, lifted

    def proof_gen(self, st, crs):
        """Prover::proof_gen (proofgen.rs:30-427) -> Transcript."""
        c, ctx, ch = self.constants, self.ctx, self.verifier.challenges
        alive = []
        cch = _chal(ch, alive)
        pi, pi2, omega, alpha, beta, cc = alive
        phi, a, b = st.phi_k[0], st.a_k[0], st.b_k[0]
        cst = _lib.CState(_p(phi), _p(a), _p(b))
        out = {
            "u_1": np.zeros((c.KAPPA_1, D), np.uint32), "projection_int": np.zeros(JL_ROWS, np.int64),
            "projection": np.zeros(JL_ROWS, np.uint32), "b_prime_prime": np.zeros(D, np.uint32),
            "u_2": np.zeros((c.KAPPA_2, D), np.uint32), "z": np.zeros((c.N, D), np.uint32),
            "t": np.zeros((c.R, c.KAPPA, D), np.uint32), "g": np.zeros((c.R, c.R, D), np.uint32),
            "h": np.zeros((c.R, c.R, D), np.uint32), "phi_final": np.zeros((c.R, c.N, D), np.uint32),
        }
        tr = _lib.CTranscript(_p(out["u_1"]), 0, _p(out["projection_int"]), _p(out["projection"]), _p(out["b_prime_prime"]),
                              _p(out["u_2"]), _p(out["z"]), _p(out["t"]), _p(out["g"]), _p(out["h"]), _p(out["phi_final"]), 0)
        S = self.witness
        seedbuf = _seed_buf(crs.base_seed)
        import time as _time
        t0 = _time.perf_counter()
        rc = ctx.L.lab_prove(ctx._h, C.byref(c), _p(seedbuf), _p(S), C.byref(cst), C.byref(cch), C.byref(tr))
        ctx.last_prove_seconds = _time.perf_counter() - t0           # the C call alone (host buffers in, transcript out), without this method's marshalling
        if rc == 1:
            raise LabError(rc, "failed JL...")                                     # proofgen.rs:176
        ctx._ck(rc)
        pi_acc = pi[tr.jl_attempt] if pi is not None else unpack_pi(pi2[tr.jl_attempt])
        return Transcript(u_1=out["u_1"], pi_accepted=pi_acc, jl_attempt=tr.jl_attempt,
                          projection_int=out["projection_int"], projection=out["projection"],
                          psi=[[int(ch["psi"])]], omega=[omega], b_prime_prime=[out["b_prime_prime"]],
                          alpha=[alpha], beta=[beta], u_2=out["u_2"], c=cc, z=out["z"], t_i_all=out["t"],
                          g_mat=out["g"], h_mat=out["h"], phi_final=out["phi_final"], norm_sum=int(tr.norm_sum))


def generate_witness(constants, seed=synth.SEED):
    """generate_witness (proofgen.rs:460-518), seeded."""
    return synth.generate_witness(constants.N, constants.R, constants.BETA_BOUND, seed)
