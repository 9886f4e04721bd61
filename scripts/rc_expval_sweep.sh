for v in 16 32 64 128 256 512 1024; do
  echo "## FASTPAULI_RC_EXPVAL_CTAS_PER_SM=$v"
  for c in local2 local3 local4 span3; do FASTPAULI_RC_EXPVAL_CTAS_PER_SM=$v python scripts/run_case.py $c --iters 10 2>&1 | grep -v Warn; done
done
