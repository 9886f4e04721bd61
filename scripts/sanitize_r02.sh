#!/bin/bash
# compute-sanitizer over the round-2 kernel families (run on a GPU box)
set -x
python tests/sanitize_case_r02.py 2>&1 | tail -15
compute-sanitizer --tool memcheck python tests/sanitize_case_r02.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|error|ok:|passed" | head -30
compute-sanitizer --tool racecheck python tests/sanitize_case_r02.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard|ok:|passed" | head -30
