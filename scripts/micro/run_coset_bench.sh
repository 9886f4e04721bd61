#!/bin/bash
cd "$(dirname "$0")"
run() { echo "## $*"; timeout 60 ./coset_bench "$@" | grep -v "pass:" || echo "FAILED/timeout: $*"; }
run cfg3 16 1024 8 4 4 148 1
run chain 20 64 0 4 4 148 1
run few16 20 64 0 4 4 148 1
