timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py tests/test_gpu_fc.py -x -q 2>&1 | tail -15
timeout 120 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2b_bench_v2.json 2> gpurun_out/r2b_bench_v2.err; echo "bench v2 rc=$?"
ENFLOW_FWD_V1=1 timeout 120 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2b_bench_v1.json 2> gpurun_out/r2b_bench_v1.err; echo "bench v1 rc=$?"
python - <<'PY'
import json
for v in ('v2','v1'):
    try:
        d=json.load(open(f'gpurun_out/r2b_bench_{v}.json'))
        print(v, d['ms_per_step'], d['kernel_ms_per_step']['edge_fwd'], d['kernel_ms_per_step']['edge_bwd'])
    except Exception as e: print(v, 'failed', e)
PY
