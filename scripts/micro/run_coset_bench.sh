#!/bin/bash
cd "$(dirname "$0")"
B=./coset_bench
run() { echo "## $*"; timeout 20 $B "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
# case n B ctPer(-1 = DSM, all tiles) log_twc nbuf grid/minb pstr
run few 20 64 -1 4 1 2 1
run few 20 64 -1 3 2 2 1
run few 20 64 -1 3 1 2 1
run few4 20 64 -1 4 1 2 1
run rand 20 64 -1 4 1 2 1
run rand 20 64 -1 3 2 2 1
run few 20 256 -1 4 1 2 1
run few 20 256 -1 3 2 2 1
