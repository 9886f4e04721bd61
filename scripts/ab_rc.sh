#!/bin/bash
cd "$(dirname "$0")/.."
for c in span1 span2 span3 span4 local2 local3 local4; do
  echo -n "rcoset "; python scripts/run_case.py $c --iters 20 2>&1 | tail -1
done
