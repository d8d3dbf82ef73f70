#!/bin/bash
for v in "$@"; do MASTB200_LIB=$v python tools/bench_configs.py --only c5 --c5 100 50 50 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', d['kernel_ms'], round(d['elements_per_s']/1e6,2))"; done
