#!/bin/bash
cd "$(dirname "$0")"
run() { echo "## $*"; timeout 60 ./coset_bench "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
# case n B chunkTiles log_twc nbuf(4 = K3g) grid cmeta
run few 12 64 0 4 4 148 1
run few16 20 64 0 4 4 148 1
run chain 20 64 0 4 4 148 1
run rand 20 64 0 4 4 148 1
run cfg3 16 1024 8 4 4 148 0
run cfg3 16 1024 8 4 4 148 1
run cfg3 16 1024 4 4 4 148 1
run cfg3 16 1024 16 4 4 148 1
run cfg3 16 128 1 4 4 148 1
run cfg3 16 128 2 4 4 148 1
