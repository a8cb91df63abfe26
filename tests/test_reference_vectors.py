"""Pins the CPU oracle against outputs of the UNMODIFIED reference crate.

The reference is Rust and this image has no cargo/rustc, so the reference cannot be run here ("parity unpinned").
labrador-snark_b200/rust/reference-vectors/examples/gen_vectors.rs is a 100-line program that links the reference crate and
prints a JSON object; anyone with a Rust toolchain runs

    cargo run --release --example gen_vectors > reference_actual.json
    LAB_REFERENCE_ACTUAL=reference_actual.json python -m pytest tests/test_reference_vectors.py -q

and this test names the first value on which the oracle and the reference disagree (or passes: parity pinned).  Without that
file the value test is skipped; what always runs is the self-consistency of the committed expectation (tests/golden/
reference_expect.json is what the oracle produces today, and the relations between overlapping CRS regions that the generator
also checks on the reference side hold in the oracle)."""
import json
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_expect as mre  # noqa: E402

EXPECT = os.path.join(HERE, "golden", "reference_expect.json")
ACTUAL = os.environ.get("LAB_REFERENCE_ACTUAL") or os.path.join(HERE, "golden", "reference_actual.json")


def diff(path, want, got, out):
    if isinstance(want, dict):
        for k, v in want.items():
            if k.endswith("_note"):
                continue
            if k not in got:
                out.append(f"{path}/{k}: missing in the reference output")
            else:
                diff(f"{path}/{k}", v, got[k], out)
    elif isinstance(want, list):
        if len(want) != len(got):
            out.append(f"{path}: length {len(got)} != {len(want)}")
            return
        for i, (a, b) in enumerate(zip(want, got)):
            diff(f"{path}[{i}]", a, b, out)
    elif isinstance(want, float):
        if not math.isclose(want, float(got), rel_tol=1e-12, abs_tol=0.0):
            out.append(f"{path}: reference {got!r} != oracle {want!r}")
    elif want != got:
        out.append(f"{path}: reference {got!r} != oracle {want!r}")


def test_committed_expectation_is_what_the_oracle_produces():
    assert json.load(open(EXPECT)) == json.loads(json.dumps(mre.expected()))


def test_overlapping_crs_regions_in_the_oracle(orc):
    """structs.rs:74-88: consecutive B_ik start KAPPA_1 * KAPPA counters apart but a row is KAPPA * D counters long, so at (2,2)
    row 0 of B_01 is row 2 of B_00; the reference-side generator asserts the same relation on the reference."""
    e = mre.crs_expected("00" * 31 + "07")
    assert e["B_0_1_0_equals_B_0_0_2"] is True
    c, _ = orc.constants(2, 2)
    seed = bytes(range(32))
    assert np.array_equal(orc.fetch_B_ik_row(c, seed, 1, 1, 0), orc.fetch_B_ik_row(c, seed, 1, 0, 2))
    # D overlaps C (structs.rs:116-144): D_00,k starts R (R + 1) / 2 vectors after C_00,0, i.e. at C's pair index 3 / T_1 ...
    assert orc.offset("D", c, 0, 0, 0) == orc.offset("C", c, 0, 0, 0) + 3 * c.KAPPA_2 * 64


@pytest.mark.skipif(not os.path.exists(ACTUAL), reason="no reference output supplied (LAB_REFERENCE_ACTUAL / tests/golden/reference_actual.json): "
                                                       "run rust/reference-vectors/examples/gen_vectors.rs against the reference crate")
def test_oracle_matches_the_reference_crate():
    got = json.load(open(ACTUAL))
    want = json.load(open(EXPECT))
    problems = []
    size_ref = got.get("bincode", {}).get("size_in_bytes")
    size_ora = want["bincode"].pop("size_in_bytes")
    diff("", want, got, problems)
    crs = got.get("crs")
    if not crs:
        problems.append("/crs: missing in the reference output")
    else:
        diff("/crs", mre.crs_expected(crs["base_seed"], crs["N"], crs["R"]), crs, problems)
    assert not problems, "oracle and reference disagree:\n  " + "\n  ".join(problems[:20])
    assert size_ref is not None and abs(size_ref - size_ora) <= 64, f"gzip size metric: reference {size_ref}, zlib here {size_ora}"
