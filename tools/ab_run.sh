#!/bin/bash
# quick A/B on one B200: GPU tests, then the headline step under a list of environment settings
#   tools/ab_run.sh <outdir> "<ENV=.. ENV=..>" "<ENV..>" ...     ("-" = default environment)
o=gpurun_out/$1; shift
mkdir -p $o
(timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | cut -c1-300) > $o/gputests.log 2>&1
grep -E "passed|failed|Error" $o/gputests.log | tail -5
i=0
for envs in "$@"; do
  i=$((i+1))
  [ "$envs" = "-" ] && envs=""
  for rep in 1 2; do
    (env $envs timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-breakdown --skip-cpu-baseline $BENCH_ARGS) > $o/ab_${i}_$rep.json 2> $o/ab_${i}_$rep.err
    python - <<PY
import json
try:
    d = json.load(open("$o/ab_${i}_$rep.json"))
    print("[$envs] rep $rep: ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches_per_step"])
except Exception as e:
    print("[$envs] failed", e, open("$o/ab_${i}_$rep.err").read()[-800:])
PY
  done
done
