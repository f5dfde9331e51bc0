#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_tc_conv_layers_against_torch tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores tests/test_gpu_round2.py::test_regression_net_bf16x3 \
         tests/test_gpu_parity.py::test_no_out_of_bounds_writes; do run $t; done
for v in "FLD_X=1"; do
echo "---- bench bf16 $v"
env $v timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --no-sub > $OUT/bench_px8.json 2> $OUT/bench_px8.err; echo "rc=$?"; tail -3 $OUT/bench_px8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_px8.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'layers', d['roofline']['layer_ms'], 'cnn', round(d['roofline_cnn']['frac_burst'],3))
PY
done
