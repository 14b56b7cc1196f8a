"""prove() of a config, then again from the resident trace: same bytes expected.  Usage: python scripts/reuse_check.py CONFIG"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genstark_b200 import workloads
from genstark_b200.stark import Stark

air, opts, a, inputs, seed, desc = workloads.config(sys.argv[1] if len(sys.argv) > 1 else '5')
st = Stark(air, opts)
p0 = st.prove_bytes(a, inputs, seed)
p1 = st.prove_bytes(a, inputs, seed)
print('second prove equal:', p0 == p1)
for i in range(3):
    try:
        p2 = st.prove_bytes(a, inputs, seed, _reuse_resident_trace=True)
        print('resident prove equal:', p2 == p0)
    except Exception as e:
        print('resident prove FAILED:', e)
