#!/usr/bin/env python3
"""Small workloads for ncu captures of the kernels other than the inner commitment:
  mv   two (8,8) proofs  -> k_crs_matvec (outer commitment u_1: the kernel behind the reference's own benchmark sizes)
  ntt  forward transform of 2^22 polynomials, three times -> k_ntt_fwd_regs"""
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
ctx.close()
