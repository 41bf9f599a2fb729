#!/bin/bash
# Round 2, GPU call C (1 GPU): two-steps-per-pass kernel — parity suite + A/B rates
OUT=gpurun_out/r02c
mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log
for dt in f32 f64; do for col in bgk trt regularized kbc; do for s2 in 1 0; do
  CHEMSIM_LBM_STEP2=$s2 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --collision $col --dtype $dt > $OUT/bench_${col}_${dt}_step2_$s2.json 2>> $OUT/bench.err
done; done; done
python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --workload config3 > $OUT/bench_config3_f32_step2_1.json 2>> $OUT/bench.err
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', 'launches', d['gpu_launches'], 'reps', d['reps'], 'drift', d['run']['mass_drift_rel'], d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -5 $OUT/bench.err
ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f32 \
    python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras > $OUT/ncu_step2.log 2>&1
