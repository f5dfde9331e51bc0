#!/bin/bash
# round-2 GPU call: new align kernel + ABI v2 tests, align microbench variants
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_align_c4_full_size tests/test_gpu_parity.py::test_align_68_to_5_and_edge_cases tests/test_gpu_parity.py::test_warp_affine_golden \
         tests/test_gpu_round2.py::test_align_tile_kernel_edges tests/test_gpu_round2.py::test_warp_affine_caller_matrices_with_shear \
         tests/test_gpu_round2.py::test_align_tile_equals_generic_kernel_at_c4_scale tests/test_properties.py; do run $t; done
echo "---- align microbench"
for v in "" "FLD_ALIGN_TILE_OFF=1" "FLD_ALIGN_YSPLIT=1" "FLD_ALIGN_YSPLIT=2" "FLD_ALIGN_YSPLIT=4"; do
  echo "== $v"; env $v timeout 300 python tools/bench_kernels.py align 2>&1 | tail -2
done
echo "---- rest of round-2 tests"
for t in tests/test_gpu_round2.py::test_decode_scratch_is_per_call_and_stream_safe tests/test_gpu_round2.py::test_capture_keeps_buffers_and_invalidates \
         tests/test_gpu_round2.py::test_video_predict_with_fake_capture tests/test_gpu_round2.py::test_c2_batch256_against_oracle \
         tests/test_gpu_round2.py::test_c3_224_against_oracle tests/test_gpu_round2.py::test_decoded_landmarks_every_encoder_bf16 \
         tests/test_gpu_parity.py::test_heatmap_xy_golden_and_oracle tests/test_gpu_parity.py::test_fcn8_fused_soft_centroid \
         tests/test_gpu_parity.py::test_pipeline_chunks_lanes_and_graph_bit_identical tests/test_gpu_parity.py::test_pipeline_end_to_end; do run $t; done
