#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1
timeout 600 python tools/prof_c3.py 64 > $OUT/prof_c3_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:deconv_gemm -s 5 -c 1 -f -o $OUT/ncu_up8_tc python tools/prof_c3.py 64 > $OUT/ncu_up8_tc.log 2>&1; echo "ncu tc rc=$?"
FLD_TC_DECONV_WALK=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:deconv_gemm -s 5 -c 1 -f -o $OUT/ncu_up8_walk python tools/prof_c3.py 64 > $OUT/ncu_up8_walk.log 2>&1; echo "ncu walk rc=$?"
