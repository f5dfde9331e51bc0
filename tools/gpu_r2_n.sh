#!/bin/bash
set -u
mkdir -p gpurun_out
python face-landmark-detector_b200/build.py > gpurun_out/build.log 2>&1
timeout 600 python tools/trace_c3.py 256 2>&1 | tail -30
