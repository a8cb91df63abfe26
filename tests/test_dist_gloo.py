"""Multi-rank path on CPU: world_size 2 over gloo.  Each rank computes its shard with the oracle (standing in for
the GPU kernels, which tests/test_gpu_parity.py checks against the same oracle) and the ranks combine exactly the
way bench.py does on NCCL: int64 all-reduce of the JL partials and of z followed by mod q, all-gather of g tiles,
row-sharded T.  The combined result must equal the single-rank result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, R, q):
    for p in (os.path.join(ROOT, "labrador-snark_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle
    from labrador_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c, _ = oracle.constants(N, R)
        seed = bytes(range(32))
        S = synth.uniform_witness(N, R, 3)
        pi = synth.sample_pi(N, R, 3, 0)
        ch = synth.prg_zq(3, 10, R * 64).reshape(R, 64)
        pl = shard.plan(c.KAPPA, R, world, rank)
        i0, ni = pl["i0"], pl["ni"]

        def allreduce(a):
            t = torch.from_numpy(a)
            dist.all_reduce(t)
            return t.numpy()

        # JL: partial over this rank's witness vectors (zero out the others), int64 all-reduce, mod q
        Sz = np.zeros_like(S); Sz[i0:i0 + ni] = S[i0:i0 + ni]
        p, p_mod = shard.combine_jl(oracle.jl_project(c, Sz, pi), allreduce)
        # z: partial over this rank's witness vectors
        z = shard.combine_z(oracle.amortize_z(c, Sz, ch), allreduce)
        # g: rows [i0, i0+ni) then all-gather (equal tile sizes here)
        G_tile = torch.from_numpy(oracle.gram(c, S)[i0:i0 + ni].astype(np.int64))
        tiles = [torch.empty_like(G_tile) for _ in range(world)]
        dist.all_gather(tiles, G_tile)
        G = torch.cat(tiles).numpy().astype(np.uint32)
        # T: row shard, stays sharded; rank 0 collects only for the comparison
        T_shard = torch.from_numpy(oracle.commit_inner_rows(c, seed, S, pl["row0"], pl["nrows"], ntt=True).astype(np.int64))
        shards = [torch.empty_like(T_shard) for _ in range(world)]
        dist.all_gather(shards, T_shard)
        T = torch.cat(shards, dim=1).numpy().astype(np.uint32)
        if rank == 0:
            ok = (np.array_equal(p, oracle.jl_project(c, S, pi)) and np.array_equal(p_mod, np.mod(oracle.jl_project(c, S, pi), 8191).astype(np.uint32))
                  and np.array_equal(z, oracle.amortize_z(c, S, ch)) and np.array_equal(G, oracle.gram(c, S))
                  and np.array_equal(T, oracle.commit_inner_rows(c, seed, S, 0, c.KAPPA, ntt=True)))
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,R", [(1, 2), (2, 4)])
def test_two_rank_sharding_equals_single_rank(N, R):
    import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, R, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True
