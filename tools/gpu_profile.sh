#!/bin/bash
# GPU box: full gpu test run, bench (plain) and — only when it exits 0 — the ncu launch list and a full-set capture of
# the conv kernels of the same command; then the contract bench lines and the per-kernel microbenchmarks.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
# eager, one lane: the launch order is the layer order, so the -s/-c window below selects one whole trunk
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --dtype bf16 --no-graph --lanes 1"
$CMD > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch-list rc=$?"
$CMD > $OUT/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 15 -c 5 -o $OUT/prof_conv $CMD > $OUT/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 $OUT/ncu_full.log
python bench.py --steps 20 --warmup 3 --dtype bf16 > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?"; cat $OUT/bench_bf16.json | cut -c1-300
python bench.py --steps 5 --warmup 3 --dtype fp32 --no-cpu > $OUT/bench_fp32.json 2> $OUT/bench_fp32.err; echo "bench fp32 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
python tools/bench_kernels.py > $OUT/kernels.jsonl 2> $OUT/kernels.err; echo "kernels rc=$?"
python tools/sweep_batch.py bfloat16 > $OUT/sweep_bf16.jsonl 2> $OUT/sweep.err; echo "sweep rc=$?"
