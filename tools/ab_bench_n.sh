#!/bin/bash
# multi-GPU A/B of the headline step: tools/ab_bench_n.sh <N> <outdir> ["ENV=val ..." ...]
n=$1; out=$2; shift 2
mkdir -p "$out"
port=29500
for v in "$@"; do
  envs=""; flags=""
  if [ "$v" != "-" ]; then
    for tok in $v; do case "$tok" in --*) flags="$flags $tok";; *=*) envs="$envs $tok";; *) flags="$flags $tok";; esac; done
  fi
  port=$((port+1))
  name=$(echo "$v" | tr ' =/' '___')
  ms=$(env $envs timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-breakdown --skip-cpu-baseline $flags 2> "$out/n${n}_$name.err" | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms  e2e %.3f ms  %s  check=%s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['gradient_exchange'], d.get('gradient_average_check',{}).get('max_rel_err_vs_allgather_mean')))")
  echo "N=$n $v : $ms" | tee -a "$out/ab.txt"
done
