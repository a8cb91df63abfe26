#!/usr/bin/env python3
"""Small workloads for ncu captures of the kernels other than the inner commitment:
  mv   two (8,8) proofs  -> k_crs_matvec (outer commitment u_1: the kernel behind the reference's own benchmark sizes)
  ntt  forward transform of 2^22 polynomials, three times -> k_ntt_fwd_regs
  mul  fused product of 2^22 polynomial pairs, three times -> k_polymul_regs
  phi  phi'' at (N, R) = (1024, 16) through lab_aggregate_phi2, twice -> k_piT_omega2"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import labrador_b200 as lb  # noqa: E402
from labrador_b200 import synth  # noqa: E402

ctx = lb.Context(0)
if "mv" in sys.argv[1:]:
    N = R = 8
    c = lb.RuntimeConstants.new(N, R)
    S = synth.generate_witness(N, R, c.BETA_BOUND, synth.SEED)
    st = lb.State.new(S, c, synth.SEED, ctx)
    ver = lb.Verifier.new(st.b_prime_k, c, seed=synth.SEED, n_attempts=6)
    prover = lb.Prover.new(S, ver, c, ctx)
    crs = lb.CRS.from_seed(c, bytes(range(32)), ctx)
    for _ in range(2):
        prover.proof_gen(st, crs)
if "ntt" in sys.argv[1:]:
    n = 1 << 22
    da, dc = ctx.malloc(n * 256), ctx.malloc(n * 256)
    ctx.synth_zq_dev(synth.SEED, 20, 0, n * 64, da)
    for _ in range(3):
        ctx.ntt_fwd_batch_dev(da, dc, n)
    ctx.sync()
if "mul" in sys.argv[1:]:
    n = 1 << 22
    da, db, dc = ctx.malloc(n * 256), ctx.malloc(n * 256), ctx.malloc(n * 256)
    ctx.synth_zq_dev(synth.SEED, 20, 0, n * 64, da)
    ctx.synth_zq_dev(synth.SEED, 21, 0, n * 64, db)
    for _ in range(3):
        ctx.polymul_batch_dev(da, db, dc, n)
    ctx.sync()
if "phi" in sys.argv[1:]:
    import numpy as np
    n2, r2 = 1024, 16
    c2 = lb.RuntimeConstants.new(n2, r2, allow_degenerate=True)
    d2 = ctx.malloc(r2 * 256 * n2 * 64 // 4)
    ctx.synth_pi2_dev(synth.SEED, 1, 0, r2 * 256 * n2 * 64, d2)
    pi2 = np.empty((r2, 256, n2 * 4), np.uint32)
    ctx.d2h(pi2, d2); ctx.sync()
    phi = synth.prg_zq(3, 4, r2 * n2 * 64).reshape(r2, n2, 64)
    for _ in range(2):
        ctx.aggregate_phi2(c2, phi, pi2, 77, synth.prg_zq(3, 7, 256))
ctx.close()
