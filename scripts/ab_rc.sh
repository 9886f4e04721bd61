#!/bin/bash
cd "$(dirname "$0")/.."
for d in 1 0; do
for c in span4 local4 span5 local5; do
  echo -n "dcoset=$d "; FASTPAULI_DCOSET=$d python scripts/run_case.py $c --iters 20 2>&1 | tail -1
done
done
