#!/bin/bash
# Round 2, GPU call Q (1 GPU): final evidence on the final build (f32 tiles 8 rows tall, packed additions, L2 prefetch
# a quarter wave ahead): parity suite, the driver's bench command, ncu launch list and full profile of the same command
OUT=gpurun_out/r02q
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log | head -2
python bench.py --steps 20 --warmup 5 > $OUT/bench_driver_cmd.json 2> $OUT/bench_driver_cmd.err
python - $OUT/bench_driver_cmd.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('driver cmd', round(d['value'],2), 'GLUPS', d['batch_ms'], d['roofline']['frac'], d['clocks'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu --no-extras > $OUT/launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f32_final \
  python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras > $OUT/ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f64_final \
  python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras --dtype f64 > $OUT/ncu64.log 2>&1
for spec in bgk:f64 regularized:f32 trt:f32; do col=${spec%%:*}; dt=${spec##*:}
  python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --collision $col --dtype $dt > $OUT/bench_${col}_${dt}.json 2>> $OUT/bench.err
  python -c "import json; d=json.loads(open('$OUT/bench_${col}_${dt}.json').read().strip().splitlines()[-1]); print('$col $dt', round(d['value'],2), d['clocks'])"
done
ls $OUT | wc -l
