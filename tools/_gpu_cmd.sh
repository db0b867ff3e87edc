timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print(d['ms_per_step'], d['kernel_ms_per_step'])
PY
