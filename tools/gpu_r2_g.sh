#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_tc_conv_layers_against_torch tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores; do run $t; done
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --no-sub > $OUT/bench_px8.json 2> $OUT/bench_px8.err; echo "rc=$?"; tail -3 $OUT/bench_px8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_px8.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'layers', d['roofline']['layer_ms'], 'cnn', round(d['roofline_cnn']['frac_burst'],3))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_px8 -c 1 -s 2 -f -o $OUT/ncu_px8 python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-graph --lanes 1 > $OUT/ncu_px8.log 2>&1; echo "ncu rc=$?"
