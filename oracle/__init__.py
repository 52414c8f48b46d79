"""CPU oracle for the sameold receiver path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this package.
Nothing under sameold_b200/ does.
"""
from .pyoracle import (  # noqa: F401
    Oracle,
    OracleConfig,
    OracleEvent,
    build_oracle,
    decode_batch_events,
    synth_cpu,
    agc_block_model,
    BATCH_EVENT_DTYPE,
    load_golden_recording,
    GOLDEN_DIR,
    EV_NAMES,
)
