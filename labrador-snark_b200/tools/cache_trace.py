"""(N, N) proof with the CRS cache: fill, verify, proof again -- wall times, and with LAB_TRACE=1 the stage timestamps of each call."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import labrador_b200 as lb
from labrador_b200 import synth

N = R = int(os.environ.get("TP_N", "32"))
ctx = lb.Context(0)
c = lb.RuntimeConstants.new(N, R)
S = synth.generate_witness(N, R, c.BETA_BOUND, synth.SEED)
st = lb.State.new(S, c, synth.SEED, ctx)
ver = lb.Verifier.new(st.b_prime_k, c, seed=synth.SEED, n_attempts=6)
prover = lb.Prover.new(S, ver, c, ctx)
crs = lb.CRS.from_seed(c, bytes(range(32)), ctx)
ctx.crs_cache_configure(150 << 30)
for i in range(3):
    t0 = time.perf_counter(); tr = prover.proof_gen(st, crs); print("prove", i, round((time.perf_counter() - t0) * 1e3, 3), "ms", ctx.crs_cache_stats(), flush=True)
    t0 = time.perf_counter(); ok = ctx.verify(c, bytes(range(32)), st.phi_k[0], st.a_k[0], st.b_k[0], ver.challenges, tr.as_oracle_dict())
    print("verify", i, round((time.perf_counter() - t0) * 1e3, 3), "ms", ok[:2], flush=True)
