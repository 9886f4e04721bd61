#!/bin/bash
# GPU sweep for scripts/micro/tma_gather.cu (design data for the TMA-fed coset kernel); every run bounded by timeout
cd "$(dirname "$0")"
B=./tma_gather
run() { timeout 90 $B "$@" || echo "FAILED/timeout: $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
# mode W lanes G ctas/SM stages box_rows
for W in 256 512 1024; do run 0 $W 32 0 1 2 1; done
# which box encoding does gather4 want?
run 1 256 32 0 1 3 1
run 1 256 32 0 1 3 4
for W in 128 256 512; do for L in 1 4 32; do run 1 $W $L 0 1 3 1; done; done
run 1 256 32 0 1 2 1
run 1 1024 32 0 1 2 1
run 1 256 32 0 2 3 1
for W in 256 512; do run 2 $W 32 0 1 3 1; run 2 $W 32 0 1 2 1; done
for G in 0 4 8 12 16; do run 4 256 32 $G 1 2 1; done
run 4 256 4 8 1 2 1
for G in 0 4 8 12 16; do run 5 256 32 $G 1 2 1; done
