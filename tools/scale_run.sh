#!/bin/bash
# Multi-GPU bench points on one box:  gpurun --gpus N -- 'bash tools/scale_run.sh N "c3 c2" [extra bench args]'
# One JSON line per (config, N) under gpurun_out/r2_scale_<config>_n<N>.json
N=$1; shift
CONFIGS=$1; shift
mkdir -p gpurun_out
for c in $CONFIGS; do
  extra=""
  if [ "$c" = "c4full" ]; then c=c4; extra="--batch 125000 --steps 5"; tag=c4_1M; else tag=$c; fi
  out=gpurun_out/r2_scale_${tag}_n${N}.json
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --config $c --steps 10 --no-cpu-baseline $extra "$@" > $out 2> ${out%.json}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --config $c --steps 10 --no-cpu-baseline $extra "$@" > $out 2> ${out%.json}.err
  fi
  echo "$tag n=$N rc=$? $(python -c "import json,sys; d=json.loads([l for l in open('$out') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d['launch_mode'][:60])" 2>&1 | tail -1)"
done
