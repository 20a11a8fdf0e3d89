#!/bin/bash
# fp32-mode bring-up on one B200: the new parity tests first (stop at the first failure of a file, keep going across files)
o=gpurun_out/$1
mkdir -p $o
(timeout 600 python -m pytest tests/test_gpu_fp32_mode.py -q 2>&1 | grep -E 'Error|assert|FAILED|passed|failed|rel L2|max-norm' | cut -c1-300 | tail -60) > $o/fp32_tests.log 2>&1
tail -15 $o/fp32_tests.log
