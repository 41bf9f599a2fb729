#!/bin/bash
# Round 2, GPU call I (1 GPU): the TMA bulk-copy A/B (single-step kernel, LDG vs cp.async.bulk loads)
OUT=gpurun_out/r02i
mkdir -p $OUT
export CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm_bulk.so CHEMSIM_LBM_STEP2=0
CHEMSIM_LBM_BULK=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_4096_against_oracle or fixed_point or config1_stable or other_collision" > $OUT/pytest_bulk.log 2>&1
tail -3 $OUT/pytest_bulk.log
for rep in 1 2; do for dt in f32 f64; do for b in 0 1; do
  CHEMSIM_LBM_BULK=$b python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --dtype $dt > $OUT/bench_${dt}_bulk${b}_r$rep.json 2>> $OUT/bench.err
  python - $OUT/bench_${dt}_bulk${b}_r$rep.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', d['clocks'])
PY
done; done; done
for dt in f32 f64; do
CHEMSIM_LBM_BULK=0 ncu --set full --clock-control none --import-source on -k regex:step_vec_kernel -s 30 -c 1 -f -o $OUT/prof_ldg_$dt \
    python bench.py --steps 20 --warmup 5 --reps 1 --no-cpu --no-extras --dtype $dt > $OUT/ncu_ldg_$dt.log 2>&1
CHEMSIM_LBM_BULK=1 ncu --set full --clock-control none --import-source on -k regex:step_bulk_kernel -s 30 -c 1 -f -o $OUT/prof_bulk_$dt \
    python bench.py --steps 20 --warmup 5 --reps 1 --no-cpu --no-extras --dtype $dt > $OUT/ncu_bulk_$dt.log 2>&1
done
ls -la $OUT | head -30
