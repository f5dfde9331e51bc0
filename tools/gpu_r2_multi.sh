#!/bin/bash
# multi-GPU call: torchrun bench at N GPUs + the multi-GPU tests
set -u
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$N.txt 2>&1
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu -rs > $OUT/t${N}_$name.log 2>&1; rc=$?; echo "$rc $1"; tail -3 $OUT/t${N}_$name.log; }
run tests/test_gpu_parity.py::test_multi_gpu_shards_bit_identical
run tests/test_gpu_round2.py::test_multi_gpu_pipeline_gathers_shards
timeout 900 python example.py > $OUT/example.log 2>&1 || (cd face-landmark-detector_b200 && timeout 900 python example.py > ../$OUT/example.log 2>&1); tail -5 $OUT/example.log
echo "---- bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench rc=$?"
tail -c 2500 $OUT/bench_${N}gpu.json; grep -v "^$" $OUT/bench_${N}gpu.err | tail -8
echo "---- reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tail -c 600
