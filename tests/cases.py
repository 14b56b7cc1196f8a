"""Shared workload builders for the parity tests: the BASELINE configs live in genstark_b200/workloads.py (bench.py and
scripts/ use them too)."""
from genstark_b200.workloads import *          # noqa: F401,F403
from genstark_b200.workloads import mimc, rescue, rescue_hash_control, poseidon, config      # noqa: F401
