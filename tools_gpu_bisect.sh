#!/bin/bash
T='tests/test_gpu_dit.py -q -m gpu --no-header -p no:cacheprovider -k test_dit_forward_matches_oracle'
for m in 1 2 4 3 5 6 7; do
  echo "=== LTX_PDL_MASK=$m"
  LTX_PDL_MASK=$m timeout 300 python -m pytest $T 2>&1 | grep -E "passed|failed|AssertionError" | tr '\n' ' '; echo
done
