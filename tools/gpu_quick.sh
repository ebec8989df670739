#!/bin/bash
# Quick GPU check: parity tests, per-feature stage timing, short bench.  gpurun -- 'bash tools/gpu_quick.sh TAG'
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/ablate.py 2>&1 | tail -12
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','stage_ms','image_gbs_frac_of_hbm_peak')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
