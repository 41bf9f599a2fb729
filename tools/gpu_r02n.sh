#!/bin/bash
# Round 2, GPU call N (1 GPU): L2 prefetch of the next wave's tile in the two-step kernels (CHEMSIM_LBM_PREFETCH),
# tile height 8 vs 16 with it, packed multiplications (pm) — A/B, then the parity suite on each candidate
OUT=gpurun_out/r02n
mkdir -p $OUT
bench() {   # bench <tag> <lib-variant|base> <collision> <dtype> <steps> <warmup> [env...]
  tag=$1; v=$2; col=$3; dt=$4; st=$5; wu=$6; shift 6
  lib=$PWD/chemsim_b200/libchemsim_lbm.so; [ $v != base ] && lib=$PWD/chemsim_b200/libchemsim_lbm_$v.so
  env CHEMSIM_LBM_LIB=$lib "$@" python bench.py --steps $st --warmup $wu --no-extras --no-cpu --collision $col --dtype $dt \
      > $OUT/bench_${tag}.json 2>> $OUT/bench.err
  python - $OUT/bench_${tag}.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', d['run']['kernel'], d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
for pf in 0 -1 -2 148; do bench bgk_base_pf$pf base bgk f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 -1 -2; do bench bgk_ty8_pf$pf s2ty8 bgk f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 -1; do bench bgk_pm_pf$pf pm bgk f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 -1; do bench reg_base_pf$pf base regularized f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 -1; do bench reg_pm_pf$pf pm regularized f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 -1; do bench bgk64_base_pf$pf base bgk f64 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
# the driver's batch length
for pf in 0 -1; do bench drv_base_pf$pf base bgk f32 20 5 CHEMSIM_LBM_PREFETCH=$pf; bench drv_ty8_pf$pf s2ty8 bgk f32 20 5 CHEMSIM_LBM_PREFETCH=$pf; done
tail -3 $OUT/bench.err
CHEMSIM_LBM_PREFETCH=-1 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_base_pf.log 2>&1; tail -2 $OUT/pytest_base_pf.log
CHEMSIM_LBM_PREFETCH=-1 CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm_s2ty8.so timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_ty8_pf.log 2>&1; tail -2 $OUT/pytest_ty8_pf.log
CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm_pm.so timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_pm.log 2>&1; tail -2 $OUT/pytest_pm.log
CHEMSIM_LBM_PREFETCH=-1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f32_packed_pf \
  python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras > $OUT/ncu.log 2>&1
ls $OUT | wc -l
