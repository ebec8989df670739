#!/bin/bash
show() { grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'ms_per_step', round(d['ms_per_step'],4), 'objects', d['objects_per_step'], {k: round(v,3) for k,v in d['stage_ms'].items()})"; }
for sb in ${SEEDS:-5000 5100 5200 5300}; do ABX_SEED_BASE=$sb python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | show "seed $sb"; done
