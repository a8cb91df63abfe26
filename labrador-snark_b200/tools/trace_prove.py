import sys, time
sys.path.insert(0, "labrador-snark_b200")
import labrador_b200 as lb
from labrador_b200 import synth
ctx = lb.Context(0)
for (N, R) in ((2, 2),):
    c = lb.RuntimeConstants.new(N, R)
    S = synth.generate_witness(N, R, c.BETA_BOUND, synth.SEED)
    st = lb.State.new(S, c, synth.SEED, ctx)
    ver = lb.Verifier.new(st.b_prime_k, c, seed=synth.SEED, n_attempts=6)
    prover = lb.Prover.new(S, ver, c, ctx)
    crs = lb.CRS.from_seed(c, bytes(range(32)), ctx)
    for i in range(5):
        t0 = time.perf_counter(); tr = prover.proof_gen(st, crs); print("prove", i, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
    T, G = tr.t_i_all, tr.g_mat
    for i in range(2):
        t0 = time.perf_counter(); ctx.commit_outer_u1(c, bytes(range(32)), T, G); print("u1", i, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
