#!/bin/bash
cd "$(dirname "$0")/.."
for lnt in 7 8; do
  for c in span1 span3 span4; do
    echo -n "rcoset lnt=$lnt "; FASTPAULI_RCOSET_LOG_NT=$lnt python scripts/run_case.py $c --iters 20 2>&1 | tail -1
  done
done
