#!/bin/bash
# A/B runs of the headline step with one switch flipped at a time (quick: no breakdown, no extras, no CPU baseline).
# usage: tools/ab_bench.sh <outdir> ["ENV=val ENV2=val|--flag ..." ...]   each argument = one variant ("-" = defaults)
out=$1; shift
mkdir -p "$out"
for v in "$@"; do
  envs=""; flags=""
  if [ "$v" != "-" ]; then
    for tok in $v; do case "$tok" in --*) flags="$flags $tok";; *=*) envs="$envs $tok";; *) flags="$flags $tok";; esac; done
  fi
  name=$(echo "$v" | tr ' =/' '___')
  ms=$(env $envs timeout 150 python bench.py --steps 20 --warmup 5 --no-extras --no-breakdown --skip-cpu-baseline $flags 2> "$out/$name.err" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms  e2e %.3f ms  launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches_per_step']))")
  echo "$v : $ms" | tee -a "$out/ab.txt"
done
