#!/bin/bash
cd "$(dirname "$0")"
B=./coset_bench
run() { timeout 120 $B "$@" || echo "FAILED/timeout: $*"; }
run few 20 64 4 4
run few 20 64 2 4
run few 20 64 1 4
run few 20 64 8 3
run few 20 64 4 3
run few 20 64 1 3
run few4 20 64 4 4
run rand 20 64 4 4
run rand 20 64 8 3
run few 20 256 16 4
run few 18 64 4 4
