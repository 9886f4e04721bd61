#!/bin/bash
cd "$(dirname "$0")"
run() { echo "## $*"; timeout 60 ./coset_bench "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
run rand8 20 64 0 4 5
run rand 20 64 0 4 5
run rand16 20 64 0 4 5
run rand 16 1024 0 4 5
