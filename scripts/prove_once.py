"""One warm + N profiled proves of the bench workload (for ncu runs: keep it short)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from genstark_b200.stark import Stark

log_steps = int(os.environ.get('GS_BENCH_LOG_STEPS', '20'))
air, steps = bench.mimc_case(log_steps)
st = Stark(air, dict(bench.OPTS))
a = bench.mimc_assertions(steps)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(n):
    proof = st.prove_bytes(a, [], [3])
print('proof bytes', len(proof), 'launches', st.context.launch_count)
