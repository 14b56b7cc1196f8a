#!/bin/bash
# Mutation fuzzing of every parser of caller-supplied bytes on the host, against an AddressSanitizer + UBSan build.
#   tests/fuzz/run.sh [iterations per leg, default 4000]
# Legs: air (AIR blob -> parse + interpreter trace), proof (serialized proof -> native verifier), airverify (AIR blob -> verifier).
set -e
cd "$(dirname "$0")"
N=${1:-4000}
OUT=${TMPDIR:-/tmp}/genstark_fuzz
mkdir -p "$OUT"
g++ -O1 -g -std=c++17 -fPIC -shared -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -o "$OUT/libhostasan.so" asan_wrap.cpp -ldl
ASAN=$(gcc -print-file-name=libasan.so)
for mode in air proof airverify; do
  for air in mimc poseidon rescue; do
    [ "$mode" != air ] && [ "$air" = rescue ] && continue
    GS_FUZZ_LIB="$OUT/libhostasan.so" LD_PRELOAD="$ASAN" ASAN_OPTIONS=detect_leaks=0 python fuzz_host_asan.py $mode $air 1 $N 2>&1 | tail -4
  done
done
