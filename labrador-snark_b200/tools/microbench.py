#!/usr/bin/env python3
"""Kernel microbenchmarks on one B200 (CUDA-event timing on the ctx stream, inputs resident in HBM).
Prints one JSON object per measurement.  Used to pick kernel variants and to measure the INT32-ALU
ceiling the CRS kernels are judged against."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import labrador_b200 as lb  # noqa: E402
from labrador_b200 import synth  # noqa: E402

SEED32 = bytes(range(32))


def timeit(ctx, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop())
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ctx = lb.Context(0)
    which = sys.argv[1:] or ["ntt", "crs", "commit"]
    if "ntt" in which:
        for logn in (16, 20, 22, 24):
            n = 1 << logn
            nbytes = n * 256
            da, db, dc = ctx.malloc(nbytes), ctx.malloc(nbytes), ctx.malloc(nbytes)
            a = synth.prg_zq(1, 1, min(n, 1 << 20) * 64).reshape(-1, 64)
            reps = n // a.shape[0]
            for r in range(reps):
                ctx.h2d(da + r * a.nbytes, a)
                ctx.h2d(db + r * a.nbytes, a)
            ctx.sync()
            for name, fn, bpp in (("ntt_fwd", lambda: ctx.ntt_fwd_batch_dev(da, dc, n), 512),
                                  ("ntt_inv", lambda: ctx.ntt_inv_batch_dev(da, dc, n), 512),
                                  ("polymul", lambda: ctx.polymul_batch_dev(da, db, dc, n), 768)):
                best, med = timeit(ctx, fn)
                print(json.dumps({"kernel": name, "log2_polys": logn, "ms_best": best, "ms_median": med,
                                  "polys_per_s": n / (med * 1e-3), "GBps_algorithmic": n * bpp / (med * 1e-3) / 1e9}), flush=True)
            for d in (da, db, dc):
                ctx.free(d)
    if "crs" in which:
        for logn in (16, 20):
            n = 1 << logn
            dout = ctx.malloc(n * 256)
            best, med = timeit(ctx, lambda: ctx.crs_expand_dev(SEED32, 12345, n, dout), reps=3, warm=1)
            print(json.dumps({"kernel": "crs_expand", "log2_polys": logn, "ms_median": med, "coeffs_per_s": n * 64 / (med * 1e-3),
                              "chacha_blocks_per_s": n * 64 / (med * 1e-3)}), flush=True)
            ctx.free(dout)
    if "commit1" in which:      # one mid-size launch pair, for ncu --set full captures
        N, R, rows = 1024, 64, 148 * 4 * 6
        c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
        dS = ctx.malloc(R * N * 256)
        ctx.synth_zq_dev(synth.SEED, 1, 0, R * N * 64, dS)
        ctx.witness_load_dev(c, dS)
        dT = ctx.malloc(R * rows * 256)
        best, med = timeit(ctx, lambda: ctx.commit_inner_dev(SEED32, 0, rows, dT), reps=2, warm=1)
        print(json.dumps({"kernel": "commit_inner", "N": N, "R": R, "rows": rows, "ms_median": med,
                          "chacha_blocks_per_s": rows * N * 64 / (med * 1e-3)}), flush=True)
        ctx.free(dS); ctx.free(dT)
    if "gen" in which:          # cold large-shape commitment (generate-then-contract) on a slice of cfg 3, for ncu captures of k_gen_planes
        N, R, rows = 4096, 64, 148 * 32
        c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
        dS = ctx.malloc(R * N * 256)
        ctx.synth_zq_dev(synth.SEED, 1, 0, R * N * 64, dS)
        ctx.witness_load_dev(c, dS)
        dT = ctx.malloc(R * rows * 256)
        best, med = timeit(ctx, lambda: ctx.commit_inner_dev(SEED32, 0, rows, dT), reps=3, warm=1)
        print(json.dumps({"kernel": "commit_inner cold (k_gen_planes + k_umma_commit)", "N": N, "R": R, "rows": rows, "ms_median": med,
                          "chacha_blocks_per_s": rows * N * 64 / (med * 1e-3)}), flush=True)
        ctx.free(dS); ctx.free(dT)
    if "gen4" in which:         # a slice of cfg 4 (N = 2^16, R = 2^8): 4 passes x 4 K-segments of contraction per row chunk
        N, R, rows = 65536, 256, 4736
        c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
        dS = ctx.malloc(R * N * 256)
        ctx.synth_zq_dev(synth.SEED, 1, 0, R * N * 64, dS)
        ctx.witness_load_dev(c, dS)
        dT = ctx.malloc(R * rows * 256)
        best, med = timeit(ctx, lambda: ctx.commit_inner_dev(SEED32, 0, rows, dT), reps=2, warm=1)
        print(json.dumps({"kernel": "commit_inner cold, cfg-4 shape", "N": N, "R": R, "rows": rows, "chunk_mb": os.environ.get("LAB_GC_CHUNK_MB", "default"),
                          "ms_median": med, "chacha_blocks_per_s": rows * N * 64 / (med * 1e-3)}), flush=True)
        ctx.free(dS); ctx.free(dT)
    if "umma" in which:         # CRS-resident inner commitment on the tensor cores (lab_umma.cuh), cfg-3 shape on fewer rows
        N, R, rows = 4096, 64, 64 * 148 * 2
        c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
        dS = ctx.malloc(R * N * 256)
        ctx.synth_zq_dev(synth.SEED, 1, 0, R * N * 64, dS)
        ctx.witness_load_dev(c, dS)
        dT = ctx.malloc(R * rows * 256)
        ctx.crs_cache_configure(rows * N * 128 + (1 << 20))
        ctx.timer_start(); ctx.commit_inner_dev(SEED32, 0, rows, dT); t_fill = ctx.timer_stop()
        best, med = timeit(ctx, lambda: ctx.commit_inner_dev(SEED32, 0, rows, dT), reps=3, warm=1)
        a_bytes = rows * N * 128
        print(json.dumps({"kernel": "commit_inner from the CRS cache (k_umma_commit + build_b + finish)", "N": N, "R": R, "rows": rows, "fill_ms": t_fill,
                          "ms_median": med, "A_GBps": a_bytes / (med * 1e-3) / 1e9, "int8_TMACs_per_s": rows * N * R * 32 * 16 / (med * 1e-3) / 1e12,
                          "cache": ctx.crs_cache_stats()}), flush=True)
        ctx.crs_cache_configure(0)
        ctx.free(dS); ctx.free(dT)
    if "prove" in which:        # default-size full proofs: latency, launch count, and the u_1 stage alone
        import time
        for (N, R) in ((2, 2), (4, 4), (8, 8)):
            c = lb.RuntimeConstants.new(N, R)
            S = synth.generate_witness(N, R, c.BETA_BOUND, synth.SEED)
            st = lb.State.new(S, c, synth.SEED, ctx)
            ver = lb.Verifier.new(st.b_prime_k, c, seed=synth.SEED, n_attempts=6)
            prover = lb.Prover.new(S, ver, c, ctx)
            crs = lb.CRS.from_seed(c, SEED32, ctx)
            tr = prover.proof_gen(st, crs)
            tr = prover.proof_gen(st, crs)      # second warm-up: the scratch arena is sized after the first call
            l0 = ctx.kernel_launches
            t0 = time.perf_counter()
            reps = 10 if N <= 4 else 3
            per = []
            for _ in range(reps):
                t1 = time.perf_counter()
                tr = prover.proof_gen(st, crs)
                per.append(round((time.perf_counter() - t1) * 1e3, 3))
            dt = (time.perf_counter() - t0) / reps
            print("per-call ms", per, "jl_attempt", tr.jl_attempt, flush=True)
            launches = (ctx.kernel_launches - l0) // reps
            t0 = time.perf_counter()
            for _ in range(reps):
                ctx.commit_outer_u1(c, SEED32, tr.t_i_all, tr.g_mat)
            du1 = (time.perf_counter() - t0) / reps
            blocks = (c.R * c.T_1 * c.KAPPA_1 * c.KAPPA + c.KAPPA * c.N + (c.R * (c.R + 1) // 2) * (c.T_1 + c.T_2) * c.KAPPA_2) * 64
            t0 = time.perf_counter()
            ok = ctx.verify(c, SEED32, st.phi_k[0], st.a_k[0], st.b_k[0], ver.challenges, tr.as_oracle_dict())
            dv = time.perf_counter() - t0
            print(json.dumps({"kernel": "prove", "N": N, "R": R, "ms_per_proof": dt * 1e3, "launches_per_proof": launches, "u1_stage_ms": du1 * 1e3,
                              "chacha_blocks": blocks, "blocks_per_s_whole_proof": blocks / dt, "verify_ms": dv * 1e3, "verify_ok": ok[0]}), flush=True)
    if "batch" in which:        # BASELINE config 5 shape: independent default-size statements, one GPU
        import time
        N, R, B = 2, 2, 256
        c = lb.RuntimeConstants.new(N, R)
        S0 = synth.generate_witness(N, R, c.BETA_BOUND, synth.SEED)
        st0 = lb.State.new(S0, c, synth.SEED, ctx)
        ch0 = synth.sample_challenges(N, R, synth.SEED, 6)
        S = np.stack([S0] * B); phi = np.stack([st0.phi_k[0]] * B); a = np.stack([st0.a_k[0]] * B); b = np.stack([st0.b_k[0]] * B)
        seeds = [bytes([i % 256]) * 32 for i in range(B)]
        for shared in (False, True):
            ctx.prove_batch(c, seeds, shared, S[:8], phi[:8], a[:8], b[:8], [ch0] * 8)
            ctx.prove_batch(c, seeds, shared, S[:8], phi[:8], a[:8], b[:8], [ch0] * 8)
            t0 = time.perf_counter()
            ctx.prove_batch(c, seeds, shared, S, phi, a, b, [ch0] * B)
            dt = time.perf_counter() - t0
            print(json.dumps({"kernel": "prove_batch", "N": N, "R": R, "statements": B, "shared_crs": shared, "proofs_per_s": B / dt, "ms_per_proof": dt / B * 1e3}), flush=True)
    if "commit" in which:
        for (N, R, rows) in ((256, 2, 148 * 4 * 8), (256, 8, 148 * 4 * 8), (256, 32, 148 * 4 * 8), (256, 64, 148 * 4 * 8), (4096, 64, 148 * 4 * 4)):
            c = lb.RuntimeConstants.new(N, R, allow_degenerate=True)
            S = synth.uniform_witness(N, R)
            dS = ctx.malloc(S.nbytes)
            ctx.h2d(dS, S)
            ctx.witness_load_dev(c, dS)
            dT = ctx.malloc(R * rows * 256)
            best, med = timeit(ctx, lambda: ctx.commit_inner_dev(SEED32, 0, rows, dT), reps=3, warm=1)
            blocks = rows * N * 64
            print(json.dumps({"kernel": "commit_inner", "N": N, "R": R, "rows": rows, "ms_median": med,
                              "chacha_blocks_per_s": blocks / (med * 1e-3), "cmacs_per_s": rows * N * R * 32 / (med * 1e-3)}), flush=True)
            ctx.free(dS); ctx.free(dT)
    ctx.close()


if __name__ == "__main__":
    main()
