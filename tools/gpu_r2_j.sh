#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_round2.py::test_staged_first_layer_equals_unstaged tests/test_gpu_parity.py::test_preprocess_faces_golden_and_oracle tests/test_gpu_parity.py::test_preprocess_faces_batched_random_boxes \
         tests/test_gpu_parity.py::test_pipeline_end_to_end tests/test_gpu_parity.py::test_pipeline_chunks_lanes_and_graph_bit_identical tests/test_gpu_round2.py::test_run_host_and_host_stream_match_run_device \
         tests/test_gpu_round2.py::test_capture_keeps_buffers_and_invalidates tests/test_gpu_round2.py::test_c2_batch256_against_oracle tests/test_gpu_parity.py::test_no_out_of_bounds_writes; do run $t; done
echo "---- bench"
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --sub bf16x3,sustained > $OUT/bench_s2d.json 2> $OUT/bench_s2d.err; echo "rc=$?"; tail -3 $OUT/bench_s2d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s2d.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'layers', d['roofline']['layer_ms'], 'cnn', round(d['roofline_cnn']['frac_burst'],3), 'launches', d['gpu_launches'])
print('x3', d['sub'].get('bf16x3')); print('sustained', d['sub'].get('sustained'))
PY
