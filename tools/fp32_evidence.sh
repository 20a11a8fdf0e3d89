#!/bin/bash
# fp32 mode evidence on one B200: launch list of ONE eager step, full ncu captures of its dominant kernels
o=gpurun_out/$1
mkdir -p $o
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_fp32_step.csv python bench.py --profile-step --precision fp32 > $o/ncu_fp32.log 2>&1
python tools/summarize_launches.py $o/launches_fp32_step.csv 30 > $o/launches_fp32_step.md 2>&1
head -14 $o/launches_fp32_step.md
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tn_kernel|split3_vec|rel_score_f32|rel_grad_f32|rel_dqk_f32|attn_fwd_kernel" -s 25 -c 22 -o $o/fp32_kernels python tools/kernel_probe_fp32.py > $o/ncu_fp32_kernels.log 2>&1
tail -3 $o/ncu_fp32_kernels.log
