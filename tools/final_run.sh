#!/bin/bash
# The round's evidence run on one B200 (everything bounded by its own timeout):
#   tools/final_run.sh <outdir under gpurun_out>
o=gpurun_out/$1
mkdir -p $o
(timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-300) > $o/gputests.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > $o/smoke.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5) > $o/bench_n1.json 2> $o/bench_n1.err
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1) > $o/bench_ref.json 2> $o/bench_ref.err
(timeout 200 python tools/step_timeline.py --relation-mode index_select --out $o/timeline_dense) > $o/tl_dense.log 2>&1
(timeout 200 python tools/step_timeline.py --relation-mode banked --out $o/timeline_banked) > $o/tl_banked.log 2>&1
(timeout 120 python tools/time_banked.py) > $o/time_banked.txt 2>&1
# launch list of ONE eager step (cold-cache, serialised: compare SHARES)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_step.csv python bench.py --profile-step > $o/ncu_launches.log 2>&1
# full captures of the kernels the profiles discuss
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_kernel|gemm_nn_kernel|banked|bank_segsum" -s 8 -c 7 -o $o/kernels python tools/kernel_probe.py > $o/ncu_kernels.log 2>&1
grep -E "passed|failed" $o/gputests.log; tail -2 $o/smoke.log; head -c 400 $o/bench_n1.json; echo; head -c 300 $o/bench_ref.json; echo; tail -3 $o/ncu_kernels.log
# the score kernel with the attention tail fused in (opt-in), for the comparison in DESIGN.md 4b
GTOS_REL_FUSED_FWD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 4 -c 1 -o $o/score_fused python tools/kernel_probe.py > $o/ncu_fused.log 2>&1
tail -2 $o/ncu_fused.log
