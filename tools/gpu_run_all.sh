#!/bin/bash
# Run on the GPU box (via gpurun): every GPU test in its own process (a trapped kernel poisons only its own
# CUDA context), then smoke and short benches.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
TESTS=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" )
: > $OUT/tests_summary.txt
for t in $TESTS; do
  name=$(echo $t | sed 's/[^A-Za-z0-9_]/_/g')
  timeout 600 python -m pytest "$t" -x -q -m gpu > $OUT/test_$name.log 2>&1
  rc=$?
  echo "$rc $t" >> $OUT/tests_summary.txt
  if [ $rc -ne 0 ]; then echo "=== FAIL($rc) $t"; tail -30 $OUT/test_$name.log; fi
done
echo "---- tests summary"; cat $OUT/tests_summary.txt
echo "---- smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $OUT/smoke.log
for dt in ${BENCH_DTYPES:-bf16 fp32}; do
  echo "---- bench $dt"
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 --dtype $dt > $OUT/bench_$dt.json 2> $OUT/bench_$dt.err; echo "bench rc=$?"
  tail -c 3000 $OUT/bench_$dt.json; tail -5 $OUT/bench_$dt.err
done
