#!/bin/bash
# GPU box: full gpu test run, then bench (plain) and, only if it exits 0, the ncu launch list and a full-set
# capture of the conv kernels of the same command.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --dtype bf16"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch-list rc=$?"
$CMD > $OUT/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 15 -c 5 -o $OUT/prof_conv $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 $OUT/ncu_full.log
python bench.py --steps 20 --warmup 3 --dtype bf16 > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench rc=$?"; cat $OUT/bench_bf16.json
