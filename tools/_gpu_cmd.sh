timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --config c5 --steps 10 --no-cpu-baseline > gpurun_out/r2i_bench_c5.json 2> gpurun_out/r2i_bench_c5.err; echo "bench c5 rc=$?"
timeout 300 python bench.py --config c2 --precision bf16 --steps 10 --no-cpu-baseline > gpurun_out/r2i_bench_c2_bf16.json 2> gpurun_out/r2i_bench_c2_bf16.err; echo "bench c2 bf16 rc=$?"
python - <<'PY'
import json
for c in ('c5','c2_bf16'):
    d=json.load(open(f'gpurun_out/r2i_bench_{c}.json'))
    print(c, round(d['value']), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items() if k in ('edge_bwd','seg_cols','edge_fwd')})
PY
