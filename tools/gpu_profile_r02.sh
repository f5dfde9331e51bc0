#!/bin/bash
# round-2 profile captures (run AFTER the plain commands have exited 0): launch list, ncu --set full of the conv trunk and of the
# alignment kernel, per-kernel microbenchmarks.  Summaries: python tools/summarize_profiles.py r02
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; exit 1; }
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-graph --lanes 1"
timeout 600 $CMD > $OUT/plain_bench.json 2> $OUT/plain_bench.err; echo "plain bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_s2d|conv_halo|conv_tma" -s 10 -c 6 -f -o $OUT/prof_conv $CMD > $OUT/ncu_conv.log 2>&1; echo "conv capture rc=$?"
timeout 600 python tools/bench_kernels.py align > $OUT/plain_align.json 2>&1; echo "plain align rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:align_tile -c 1 -f -o $OUT/prof_align python tools/bench_kernels.py align > $OUT/ncu_align.log 2>&1; echo "align capture rc=$?"
timeout 900 python tools/bench_kernels.py align decode preprocess fcn > $OUT/kernels.jsonl 2> $OUT/kernels.err; echo "kernels rc=$?"; tail -2 $OUT/kernels.err
