"""labrador_b200 -- host-side harness over liblabrador_b200.so (B200-native LaBRADOR prover hot path)."""
from ._lib import Constants, D, JL_ROWS, LabError, Q, SO_PATH, SYMBOLS  # noqa: F401
from .api import (CRS, Context, Prover, RuntimeConstants, State, Transcript, Verifier,  # noqa: F401
                  default_context, generate_witness)
from . import api, fs, shard, synth  # noqa: F401
