#!/bin/bash
# Round 2, GPU call A (1 GPU): parity tests, headline bench with sub-records, per-operator rates, ncu.
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.csv
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log
( time python bench.py --steps 20 --warmup 5 ) > $OUT/bench_driver_cmd.json 2> $OUT/bench_driver_cmd.err
tail -c 600 $OUT/bench_driver_cmd.err
python bench.py --steps 200 --warmup 20 --no-extras --no-cpu > $OUT/bench_bgk_f32.json 2>> $OUT/bench.err
python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --dtype f64 > $OUT/bench_bgk_f64.json 2>> $OUT/bench.err
for col in trt regularized kbc; do
  python bench.py --steps 200 --warmup 20 --no-cpu --collision $col > $OUT/bench_${col}_f32.json 2>> $OUT/bench.err
  python bench.py --steps 100 --warmup 20 --no-cpu --collision $col --dtype f64 > $OUT/bench_${col}_f64.json 2>> $OUT/bench.err
done
python bench.py --steps 200 --warmup 20 --no-cpu --workload config3 > $OUT/bench_config3_f32.json 2>> $OUT/bench.err
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],2), d['run']['kernel'], 'reps', d['reps'], d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
# ncu: launch list of the driver's command, then one full capture per operator
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_bench_config2_f32.csv \
    python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu --no-extras > $OUT/ncu_launches.log 2>&1
for col in bgk regularized trt kbc; do
  ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 30 -c 1 -f -o $OUT/prof_${col}_f32 \
      python bench.py --steps 20 --warmup 5 --reps 1 --no-cpu --no-extras --collision $col > $OUT/ncu_${col}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 30 -c 1 -f -o $OUT/prof_bgk_f64 \
    python bench.py --steps 20 --warmup 5 --reps 1 --no-cpu --no-extras --dtype f64 > $OUT/ncu_bgk_f64.log 2>&1
ls -la $OUT
