#!/bin/bash
cd "$(dirname "$0")"
run() { echo "## $*"; timeout 60 ./coset_bench "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
runp() { echo "## prof $*"; timeout 60 ./coset_bench_prof "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
run rand8 20 64 0 4 5
run rand 20 64 0 4 5
runp rand8 20 64 0 4 5
runp rand16 20 64 0 4 5
