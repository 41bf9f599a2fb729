#!/bin/bash
# Round 2, GPU call O (1 GPU): the final build (f32 tiles 8 rows tall, packed additions, L2 prefetch half a wave ahead):
# prefetch-distance sweep, parity suite, the driver's bench command + reference arm, ncu launch list and full profile
OUT=gpurun_out/r02o
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
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log | head -2
python bench.py --steps 20 --warmup 5 > $OUT/bench_driver_cmd.json 2> $OUT/bench_driver_cmd.err; tail -c 600 $OUT/bench_driver_cmd.json | head -c 300; echo
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; head -c 300 $OUT/bench_reference_arm.json; echo
for pf in 0 64 148 222 296 444; do bench bgk_pf$pf base bgk f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 148 296; do bench drv_pf$pf base bgk f32 20 5 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 148 296; do bench bgk64_pf$pf base bgk f64 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
for pf in 0 296; do bench reg_pf$pf base regularized f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; bench trt_pf$pf base trt f32 200 20 CHEMSIM_LBM_PREFETCH=$pf; done
bench bgk_ty16_pf148 s2ty16 bgk f32 200 20 CHEMSIM_LBM_PREFETCH=148
env CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm.so python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --workload config3 > $OUT/bench_config3.json 2>> $OUT/bench.err; python -c "import json; d=json.loads(open(\"$OUT/bench_config3.json\").read().strip().splitlines()[-1]); print(\"config3\", round(d[\"value\"],2))"
tail -3 $OUT/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu --no-extras > $OUT/launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f32_final \
  python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras > $OUT/ncu.log 2>&1
ls $OUT | wc -l
