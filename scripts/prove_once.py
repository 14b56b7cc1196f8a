"""One warm + N profiled proves of a bench workload (for ncu runs: keep it short).
Usage: python scripts/prove_once.py [N] [config]   (config: a key of genstark_b200.workloads.CONFIGS, default ns)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genstark_b200 import workloads
from genstark_b200.stark import Stark

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
air, opts, a, inputs, seed, desc = workloads.config(sys.argv[2] if len(sys.argv) > 2 else 'ns')
st = Stark(air, opts)
for i in range(n):
    proof = st.prove_bytes(a, inputs, seed)
print(desc, '| proof bytes', len(proof), 'launches', st.context.launch_count)
