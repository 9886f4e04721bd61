#!/bin/bash
# Forced coset tile shapes against the heuristic's pick (20 qubits x 64 complex128 unless the case says otherwise).
for c in few20 rand20 heis20 tfim20 cfg3; do
  echo "== $c"
  python scripts/run_case.py $c --iters 5 | sed 's/^/default      : /'
  for v in 0 1 2 3 4; do
    python scripts/run_case.py $c --iters 5 --log-twc $v 2>&1 | tail -1 | sed "s/^/log_twc=$v    : /"
    FASTPAULI_COSET_VPT=8 python scripts/run_case.py $c --iters 5 --log-twc $v 2>&1 | tail -1 | sed "s/^/log_twc=$v vpt8: /"
  done
done
