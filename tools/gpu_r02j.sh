#!/bin/bash
# Round 2, GPU call J (1 GPU, final build): full parity suite, KBC launch-bound A/B, final bench lines, launch list, ncu of the two-step kernels
OUT=gpurun_out/r02j
mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log | head -2
for v in base kbc4; do for dt in f32 f64; do
  lib=$PWD/chemsim_b200/libchemsim_lbm.so; [ $v = kbc4 ] && lib=$PWD/chemsim_b200/libchemsim_lbm_kbc4.so
  CHEMSIM_LBM_LIB=$lib python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --collision kbc --dtype $dt > $OUT/bench_kbc_${dt}_$v.json 2>> $OUT/bench.err
  python - $OUT/bench_kbc_${dt}_$v.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', d['clocks'])
PY
done; done
( time python bench.py --steps 20 --warmup 5 ) > $OUT/bench_driver_cmd.json 2> $OUT/bench_driver_cmd.err
python bench.py > $OUT/bench_default.json 2>> $OUT/bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_arm.json 2>> $OUT/bench.err
for spec in bgk:f64 trt:f32 trt:f64 regularized:f32 regularized:f64; do col=${spec%%:*}; dt=${spec##*:}
  python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --collision $col --dtype $dt > $OUT/bench_${col}_${dt}.json 2>> $OUT/bench.err
done
python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --workload config3 > $OUT/bench_config3_f32.json 2>> $OUT/bench.err
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],3), 'GLUPS', d.get('run',{}).get('kernel'), 'frac', round(d.get('roofline',{}).get('frac',0),3), 'e2e', round(d['e2e']['value'],2), d.get('clocks'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/launches_bench_config2_f32.csv \
    python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu --no-extras > $OUT/ncu_launches.log 2>&1
for dt in f32 f64; do
ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_$dt \
    python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras --dtype $dt > $OUT/ncu_step2_$dt.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_regularized_f32 \
    python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras --collision regularized > $OUT/ncu_step2_reg.log 2>&1
tail -3 $OUT/bench.err
