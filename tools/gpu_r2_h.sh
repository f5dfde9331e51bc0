#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_align_c4_full_size tests/test_gpu_parity.py::test_align_68_to_5_and_edge_cases tests/test_gpu_parity.py::test_warp_affine_golden \
         tests/test_gpu_round2.py::test_align_tile_kernel_edges tests/test_gpu_round2.py::test_warp_affine_caller_matrices_with_shear \
         tests/test_gpu_round2.py::test_align_tile_equals_generic_kernel_at_c4_scale tests/test_gpu_round2.py::test_px8_first_layer_opt_in; do run $t; done
echo "---- align microbench"
for v in "" "FLD_ALIGN_YSPLIT=1" "FLD_ALIGN_RING_KB=21"; do
  echo "== $v"; env $v timeout 300 python tools/bench_kernels.py align 2>&1 | tail -1 | cut -c1-220
done
