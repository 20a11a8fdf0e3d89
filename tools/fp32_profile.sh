#!/bin/bash
# launch list (ncu, serialised, cold cache: compare SHARES) of ONE eager cfg2 step in fp32 mode
o=gpurun_out/$1
mkdir -p $o
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_fp32_step.csv python bench.py --profile-step --precision fp32 > $o/ncu_fp32.log 2>&1
python tools/summarize_launches.py $o/launches_fp32_step.csv > $o/launches_fp32_step.md 2>&1
head -40 $o/launches_fp32_step.md
