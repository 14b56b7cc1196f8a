#!/bin/bash
for env in "GS_NTT2=1 GS_NTT2_TMA=1" "GS_NTT2=1 GS_NTT2_TMA=0" "GS_NTT2=0"; do
  echo "== $env"; env $env timeout 120 python scripts/reuse_check.py 5 2>&1 | tail -5
done
echo "== config 3"; timeout 120 python scripts/reuse_check.py 3 2>&1 | tail -5
