timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
for c in c4 c1 c2; do
  timeout 300 python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/r2l_bench_$c.json 2> gpurun_out/r2l_bench_$c.err; echo "bench $c rc=$?"
done
python - <<'PY'
import json
for c in ('c4','c1','c2'):
    d=json.load(open(f'gpurun_out/r2l_bench_{c}.json'))
    print(c, round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items() if k in ('edges','node_pre','edge_fwd','node_post','run_sum')})
PY
