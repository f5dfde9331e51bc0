#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_align_c4_full_size tests/test_gpu_parity.py::test_align_68_to_5_and_edge_cases tests/test_gpu_parity.py::test_warp_affine_golden \
         tests/test_gpu_round2.py::test_align_tile_kernel_edges tests/test_gpu_round2.py::test_warp_affine_caller_matrices_with_shear \
         tests/test_gpu_round2.py::test_align_tile_equals_generic_kernel_at_c4_scale; do run $t; done
echo "---- align microbench"
for v in "" "FLD_ALIGN_YSPLIT=1" "FLD_ALIGN_YSPLIT=2" "FLD_ALIGN_YSPLIT=4"; do
  echo "== $v"; env $v timeout 300 python tools/bench_kernels.py align 2>&1 | tail -1 | cut -c1-220
done
for t in tests/test_gpu_round2.py::test_regression_net_bf16x3 tests/test_gpu_round2.py::test_c2_batch256_against_oracle \
         tests/test_gpu_round2.py::test_c3_224_against_oracle tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores \
         tests/test_gpu_parity.py::test_pipeline_chunks_lanes_and_graph_bit_identical tests/test_gpu_parity.py::test_pipeline_end_to_end; do run $t; done
echo "---- bench"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "bench rc=$?"; tail -c 6000 $OUT/bench_full.json; tail -5 $OUT/bench_full.err
env FLD_ALIGN_YSPLIT=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:align_tile -c 1 -f -o $OUT/ncu_align_v2 python tools/bench_kernels.py align > $OUT/ncu_align_v2.log 2>&1; echo "ncu rc=$?"
