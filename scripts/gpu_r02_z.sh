#!/bin/bash
# round 2, call z (1 GPU): bench.py after the config-builder edit (short line)
timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-sections --skip-anchor --skip-cg --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['config'])"
