timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
for c in c2 c1; do
  timeout 300 python bench.py --config $c --steps 20 --no-cpu-baseline > gpurun_out/r2o_bench_$c.json 2> gpurun_out/r2o_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_$c.json')); print('$c', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
