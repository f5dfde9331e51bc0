#!/bin/bash
# round-2 GPU call: x3 mode tests + ncu profile of the align tile kernel
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_round2.py::test_regression_net_bf16x3 tests/test_gpu_round2.py::test_c2_batch256_against_oracle \
         tests/test_gpu_round2.py::test_c3_224_against_oracle tests/test_gpu_round2.py::test_decoded_landmarks_every_encoder_bf16 \
         tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores tests/test_gpu_parity.py::test_tc_conv_layers_against_torch; do run $t; done
echo "---- bench x3 / bf16"
for dt in bf16x3 bf16; do timeout 600 python bench.py --steps 20 --warmup 3 --dtype $dt --no-cpu > $OUT/bench_$dt.json 2> $OUT/bench_$dt.err; echo "bench $dt rc=$?"; tail -c 1500 $OUT/bench_$dt.json; tail -3 $OUT/bench_$dt.err; done
echo "---- ncu align"
for v in "FLD_ALIGN_YSPLIT=2" "FLD_ALIGN_YSPLIT=7"; do
  env $v timeout 600 ncu --set full --import-source on --clock-control none -k regex:align_tile -c 2 -f -o $OUT/ncu_align_${v##*=} python tools/bench_kernels.py align > $OUT/ncu_align_${v##*=}.log 2>&1; echo "ncu $v rc=$?"
done
