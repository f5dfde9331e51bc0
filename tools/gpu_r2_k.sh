#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -30 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_fcn8_fused_soft_centroid tests/test_gpu_parity.py::test_full_size_properties_c3_c5 tests/test_gpu_round2.py::test_c3_224_against_oracle \
         tests/test_gpu_round2.py::test_decoded_landmarks_every_encoder_bf16 tests/test_gpu_parity.py::test_fcn8_other_class_counts_bf16 tests/test_gpu_parity.py::test_fcn8_fused_classmap tests/test_gpu_parity.py::test_fcn8_bf16; do run $t; done
for v in "FLD_X=1" "FLD_TC_DECONV_EW8=1"; do
echo "---- bench c3 $v"
env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --sub c3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "rc=$?"; tail -3 $OUT/bench_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
print(json.dumps(d['sub'].get('c3'))[:700])
PY
done
