#!/bin/bash
cd "$(dirname "$0")"
run() { echo "## $*"; timeout 60 ./coset_bench "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
# case n B ctPer(-1 = RMWPF, all tiles) log_twc nbuf minb pstr
run rand 20 64 8 3 2 2 1
run rand 20 64 -1 3 2 2 1
run few16 20 64 8 3 2 2 1
run few16 20 64 -1 3 2 2 1
run rand 20 256 -1 3 2 2 1
run rand 20 256 32 3 2 2 1
run rand 12 64 -1 3 2 2 1
