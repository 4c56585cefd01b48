#!/bin/bash
python tools/e2e_probe.py hdri_test 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -x -q -k "hdri or texture or environment or low_spp" 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --workload hdri_test --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hdri value', j['value'], 'e2e', j['e2e'])"
