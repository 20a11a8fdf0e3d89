for mode in "GTOS_SIDE_STREAM=1" "GTOS_SIDE_STREAM=0" "GTOS_GRAD_TAG=0" "GTOS_PDL=0"; do
  f=0
  for i in 1 2 3 4 5 6 7 8; do
    r=$(env $mode timeout 60 python -m pytest tests/test_gpu_modules.py -q -k "transformer_external" 2>&1 | grep -E "rel L2|passed|failed" | tr '\n' ' ' | cut -c1-160)
    case "$r" in *failed*) f=$((f+1)); echo "$mode run $i: $r";; esac
  done
  echo "$mode: $f of 8 failed"
done
