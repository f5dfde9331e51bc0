#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1
for v in 0 4 8 12; do
echo "---- bench c3 WALK=$v"
FLD_TC_DECONV_WALK=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --sub c3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "rc=$?"; tail -3 $OUT/bench_c3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3.json').read().strip().splitlines()[-1])
c=d['sub'].get('c3'); print(c['images_per_s'], c['up8_ms'], c['ms_per_batch'])
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:deconv_gemm -s 5 -c 1 -f -o $OUT/ncu_up8_tc2 python tools/prof_c3.py 64 > $OUT/ncu_up8_tc2.log 2>&1; echo "ncu tc rc=$?"
