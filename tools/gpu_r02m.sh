#!/bin/bash
# Round 2, GPU call M (1 GPU): packed f32 additions (F32x2, d2q9.cuh) — A/B against the scalar build,
# parity suite on the packed default, single-step packed variants, ncu of the packed two-step kernel
OUT=gpurun_out/r02m
mkdir -p $OUT
bench() {   # bench <tag> <lib-variant|base> <collision> [env...]
  tag=$1; v=$2; col=$3; shift 3
  lib=$PWD/chemsim_b200/libchemsim_lbm.so; [ $v != base ] && lib=$PWD/chemsim_b200/libchemsim_lbm_$v.so
  env CHEMSIM_LBM_LIB=$lib "$@" python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --collision $col --dtype f32 \
      > $OUT/bench_${tag}.json 2>> $OUT/bench.err
  python - $OUT/bench_${tag}.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', d['run']['kernel'], d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
for rep in 1 2; do
  for col in bgk regularized trt; do
    bench ${col}_packed_r$rep base $col
    bench ${col}_scalar_r$rep scalar $col
  done
done
bench bgk_packed_ty8 s2ty8 bgk
for col in kbc regularized trt; do
  bench ${col}_single_base base $col CHEMSIM_LBM_STEP2=0
  bench ${col}_single_vp3 vp3 $col CHEMSIM_LBM_STEP2=0
done
bench kbc_single_vp2 vp2 kbc CHEMSIM_LBM_STEP2=0
bench regularized_single_vp2 vp2 regularized CHEMSIM_LBM_STEP2=0
bench kbc_single_vp vp kbc CHEMSIM_LBM_STEP2=0
tail -3 $OUT/bench.err
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log | head -3
CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm_vp3.so CHEMSIM_LBM_STEP2=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "other_collision or golden or main_rs_active or full_size_4096 or ragged" > $OUT/pytest_vp3_single.log 2>&1; tail -2 $OUT/pytest_vp3_single.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f32_packed \
  python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras > $OUT/ncu.log 2>&1
ls -la $OUT | head -40
