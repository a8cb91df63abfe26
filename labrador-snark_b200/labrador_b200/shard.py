"""Sharding plan of one proof over the GPUs of a box and the combine rules.

  G1 inner commitments : by rows of A (each rank regenerates only its slice of the CRS); T stays sharded
  G2 / G7 g, h         : by rows i of the (i, j) tile grid; tiles are all-gathered
  G4 JL projection     : by witness vector i; int64 all-reduce(SUM) of the 256 partial sums, then mod q
  G9 z                 : by witness vector i; canonical partials widened to int64, all-reduce(SUM), then mod q
  batched proofs/NTT   : round-robin over ranks, no collective
All combines are exact integer sums, so the result is independent of the rank count.
"""
import numpy as np

Q = 8191


def split(total, parts, idx):
    """Contiguous balanced split: returns (start, count) of part idx."""
    base, rem = divmod(int(total), int(parts))
    return idx * base + min(idx, rem), base + (1 if idx < rem else 0)


def plan(kappa, R, world, rank):
    row0, nrows = split(kappa, world, rank)
    i0, ni = split(R, world, rank)
    return {"row0": row0, "nrows": nrows, "i0": i0, "ni": ni}


def combine_jl(partial_int64, all_reduce_sum):
    """partial_int64: this rank's exact partial projection (int64[256]); all_reduce_sum: callable doing an
    in-place SUM all-reduce over ranks.  Returns (exact projection, projection mod q)."""
    p = np.asarray(partial_int64, dtype=np.int64).copy()
    p = all_reduce_sum(p)
    return p, np.mod(p, Q).astype(np.uint32)


def combine_z(partial_canonical, all_reduce_sum):
    """partial_canonical: this rank's z partial, canonical residues.  int64 sum over ranks, then mod q."""
    z = np.asarray(partial_canonical, dtype=np.int64).copy()
    z = all_reduce_sum(z)
    return np.mod(z, Q).astype(np.uint32)


def rows_of(total, world, rank):
    """The library's own rule (lab_comm_*, shard_rows in lab_api.cu): an equal slice of `total` output rows per rank when
    the communicator size divides it, otherwise every rank computes all rows and nothing is exchanged.
    Returns (x0, nx, sharded)."""
    total, world, rank = int(total), int(world), int(rank)
    if world > 1 and total % world == 0:
        nx = total // world
        return nx * rank, nx, True
    return 0, total, False
