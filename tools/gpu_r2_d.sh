#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_tc_conv_layers_against_torch tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores tests/test_gpu_round2.py::test_regression_net_bf16x3 \
         tests/test_gpu_round2.py::test_c2_batch256_against_oracle tests/test_gpu_parity.py::test_fcn8_bf16 tests/test_gpu_encoders.py::test_fcn8_encoder_bf16 \
         tests/test_gpu_parity.py::test_no_out_of_bounds_writes tests/test_gpu_parity.py::test_pipeline_chunks_lanes_and_graph_bit_identical; do run $t; done
echo "---- bench bf16 (px8 conv1)"
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --sub bf16x3 > $OUT/bench_px8.json 2> $OUT/bench_px8.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_px8.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'layers', d['roofline']['layer_ms'], 'cnn', round(d['roofline_cnn']['frac_burst'],3), 'x3', d['sub'].get('bf16x3'))
PY
echo "---- bench bf16 (old conv1)"
FLD_C1_PX8_OFF=1 timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --no-sub > $OUT/bench_oldc1.json 2> $OUT/bench_oldc1.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_oldc1.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'layers', d['roofline']['layer_ms'])
PY
echo "---- align ring experiments"
for v in "FLD_ALIGN_RING_KB=27" "FLD_ALIGN_RING_KB=21" "FLD_ALIGN_RING_KB=16" "FLD_ALIGN_RING_KB=12" "FLD_ALIGN_RING_KB=36"; do
  echo "== $v"; env $v timeout 300 python tools/bench_kernels.py align 2>&1 | tail -1 | cut -c1-200
done
