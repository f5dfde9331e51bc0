#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_tc_conv_layers_against_torch tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores tests/test_gpu_round2.py::test_regression_net_bf16x3 \
         tests/test_gpu_round2.py::test_c2_batch256_against_oracle tests/test_gpu_parity.py::test_no_out_of_bounds_writes tests/test_gpu_parity.py::test_fcn8_bf16; do run $t; done
for v in "FLD_X=1" "FLD_C1_S2D=0"; do
echo "---- bench bf16 $v"
env $v timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --sub bf16x3 > $OUT/bench_s2d.json 2> $OUT/bench_s2d.err; echo "rc=$?"; tail -3 $OUT/bench_s2d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s2d.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'layers', d['roofline']['layer_ms'], 'cnn', round(d['roofline_cnn']['frac_burst'],3), 'x3', d['sub'].get('bf16x3'))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"widen|conv_s2d|conv_first" -c 12 --csv --log-file $OUT/launches_s2d.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-graph --lanes 1 > /dev/null 2>&1
grep -E "widen|conv_s2d|conv_first" $OUT/launches_s2d.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | head -12
