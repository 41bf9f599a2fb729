#!/bin/bash
# quick A/B on one GPU: optional parity suite, then bench lines for the listed "collision:dtype" pairs
OUT=gpurun_out/${TAG:-quick}
mkdir -p $OUT
if [ "${TESTS:-1}" = "1" ]; then ( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log | head -2; fi
for spec in ${SPECS:-bgk:f32 bgk:f64 regularized:f32 trt:f32 kbc:f32}; do
  col=${spec%%:*}; dt=${spec##*:}
  python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --collision $col --dtype $dt > $OUT/bench_${col}_${dt}.json 2>> $OUT/bench.err
  python - "$OUT/bench_${col}_${dt}.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', 'launches', d['gpu_launches'], d['run']['kernel'], d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 $OUT/bench.err
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:$NCU -s 10 -c 1 -f -o $OUT/prof python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras ${NCU_ARGS} > $OUT/ncu.log 2>&1
fi
