#!/bin/bash
# A/B of two builds of libmast_b200.so on the same box: alternating bench.py runs (kernel-only numbers)
for rep in 1 2 3; do
  for v in "$@"; do
    MASTB200_LIB=$v python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-sens 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,1), round(d['roofline']['kernel_ms'],3))"
  done
done
