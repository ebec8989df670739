#!/bin/bash
# Run on the GPU box (under gpurun): parity tests, a short bench, the ncu launch list of the same
# bench command and one `--set full` capture of the hot kernels.  TAG names the outputs.
#   gpurun --timeout 900 -- 'bash tools/gpu_profile.sh r01e [ncu]'      ("ncu": skip tests / bench / launch list)
set -u
TAG=${1:-run}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs"
if [ "$MODE" = "full" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
  echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
  python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
  echo "bench rc=$?"; cat $OUT/bench_$TAG.json
  $BENCH > $OUT/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
  echo "launch list rc=$?"
fi
# warm-up 3 steps + timed step 1 = 4 x (label_scan, stats_warp, edt_warp) launches skipped, then one step captured
$BENCH > $OUT/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'object_sweep|edt_grid|label_scan_kernel|finalize_kernel|plan_kernel|init_records|label_max' -s 28 -c 7 -f -o $OUT/prof_$TAG $BENCH > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -2 $OUT/ncu_full_$TAG.log
