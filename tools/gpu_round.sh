#!/bin/bash
# One GPU-box pass: parity tests, the default bench, the ncu launch list and the full capture of the edge kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2a'
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err; echo "bench rc=$?"
cmd="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_edge_(fwd|bwd)_tc' -s 20 -c 2 -f -o gpurun_out/${tag}_edge $cmd > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full rc=$?"
