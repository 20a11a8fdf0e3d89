#!/bin/bash
# full check on one B200: every GPU test, smoke, the default bench line (with the fp32_mode leg) - each bounded
o=gpurun_out/$1
mkdir -p $o
(timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-300) > $o/gputests.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > $o/smoke.log 2>&1
(timeout 700 python bench.py --steps 20 --warmup 5) > $o/bench_n1.json 2> $o/bench_n1.err
cp gpurun_out/fp32_mode_errors.json $o/ 2>/dev/null
grep -E "passed|failed" $o/gputests.log; tail -3 $o/smoke.log; head -c 300 $o/bench_n1.json; echo; python - <<PY
import json
try:
    d = json.load(open("$o/bench_n1.json"))
    print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "fp32", d.get("fp32_mode"))
except Exception as e:
    print("bench parse failed", e); print(open("$o/bench_n1.err").read()[-2000:])
PY
