"""Default-size proofs in a loop (first one on the ordinary path, the rest as CUDA-graph replays): wall time per proof, and the
target of `ncu --metrics gpu__time_duration.sum` launch lists of a small proof (profiles/ncu_launches_prove22_r2.csv)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import labrador_b200 as lb
from labrador_b200 import synth

ctx = lb.Context(0)
N = R = int(os.environ.get("TP_N", "2"))
reps = int(os.environ.get("TP_REPS", "6"))
c = lb.RuntimeConstants.new(N, R)
S = synth.generate_witness(N, R, c.BETA_BOUND, synth.SEED)
st = lb.State.new(S, c, synth.SEED, ctx)
ver = lb.Verifier.new(st.b_prime_k, c, seed=synth.SEED, n_attempts=6)
prover = lb.Prover.new(S, ver, c, ctx)
crs = lb.CRS.from_seed(c, bytes(range(32)), ctx)
for i in range(reps):
    t0 = time.perf_counter(); tr = prover.proof_gen(st, crs); print("prove", i, round((time.perf_counter() - t0) * 1e3, 4), "ms", flush=True)
print(ctx.graph_stats())
