#!/bin/bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python face-landmark-detector_b200/build.py > $OUT/build.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/build.log; }
run() { name=$(echo $1 | sed 's/[^A-Za-z0-9_]/_/g'); timeout 900 python -m pytest "$1" -x -q -m gpu > $OUT/t_$name.log 2>&1; rc=$?; echo "$rc $1"; if [ $rc -ne 0 ]; then tail -40 $OUT/t_$name.log; fi; }
for t in tests/test_gpu_parity.py::test_tc_conv_layers_against_torch tests/test_gpu_parity.py::test_regression_net_bf16_tensor_cores tests/test_gpu_round2.py::test_regression_net_bf16x3 \
         tests/test_gpu_round2.py::test_c2_batch256_against_oracle tests/test_gpu_parity.py::test_fcn8_bf16 tests/test_gpu_parity.py::test_no_out_of_bounds_writes; do run $t; done
echo "---- bench bf16 (px8 conv1)"
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu --sub bf16x3 > $OUT/bench_px8.json 2> $OUT/bench_px8.err; echo "rc=$?"; tail -3 $OUT/bench_px8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_px8.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'layers', d['roofline']['layer_ms'], 'cnn', round(d['roofline_cnn']['frac_burst'],3), 'x3', d['sub'].get('bf16x3'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"widen|conv_px8|resize|align_tile|regress|dense_reduce|conv_tma|conv_halo" -c 60 --csv --log-file $OUT/launches_px8.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sub --no-graph --lanes 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_px8.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
acc=collections.defaultdict(list)
for r in rows[1:]:
    try: acc[r[ki][:60]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in acc.items(): print('%-62s n=%d median %.1f us'%(k,len(v),sorted(v)[len(v)//2]/1000 if max(v)>1000 else sorted(v)[len(v)//2]))
PY
