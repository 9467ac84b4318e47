#!/bin/bash
# The CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer: builds an instrumented copy of oracle/p2de_oracle.cpp, swaps it
# in for the portable build, runs every CPU test that drives the oracle, and restores the normal build.  (This is how the dangling
# phase-timer reference fixed in round 2 was found.)      usage: tools/oracle_sanitize.sh
set -u
cd "$(dirname "$0")/.."
python -c "from oracle import oracle as o; o.build()" || exit 1
cp oracle/libp2de_oracle.so /tmp/libp2de_oracle_plain.so
trap 'cp /tmp/libp2de_oracle_plain.so oracle/libp2de_oracle.so; touch oracle/libp2de_oracle.so' EXIT
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fopenmp -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -fno-omit-frame-pointer -shared -o oracle/libp2de_oracle.so oracle/p2de_oracle.cpp || exit 1
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
python -m pytest tests/test_oracle_crosscheck.py tests/test_oracle_1d.py tests/test_oracle_bounds.py tests/test_oracle_gauss2d.py \
    tests/test_oracle_invariants.py tests/test_oracle_physics.py tests/test_golden.py -m "not gpu" -q -x -s -p no:cacheprovider
