python bench.py > gpurun_out/r2p_bench_c2.json 2> gpurun_out/r2p_bench_c2.err; echo "bench rc=$?"
cmd="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$cmd > gpurun_out/r2p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches.csv $cmd > gpurun_out/r2p_ncu1.log 2>&1
echo "ncu launches rc=$?"
for c in c3 c4 c5; do
  timeout 300 python bench.py --config $c --steps 20 --no-cpu-baseline > gpurun_out/r2p_bench_$c.json 2> gpurun_out/r2p_bench_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/r2p_bench_$c.json')); print('$c', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
