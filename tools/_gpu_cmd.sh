python bench.py > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench_c2.err; echo "bench rc=$?"
cmd="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$cmd > gpurun_out/r2j_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_launches.csv $cmd > gpurun_out/r2j_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_edge_(fwd|bwd)_tc' -s 14 -c 2 -f -o gpurun_out/r2j_edge $cmd > gpurun_out/r2j_ncu2.log 2>&1
echo "ncu full rc=$?"
