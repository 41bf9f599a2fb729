#!/bin/bash
# Round 2, GPU call T (1 GPU): phase A on horizontally adjacent cell pairs (hpair) — parity suite on that build, A/B
# against the default, streaming stores (stcs) and 10-row tiles (s2ty10) in the driver's batch length
OUT=gpurun_out/r02t
mkdir -p $OUT
CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm_hpair.so timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_hpair.log 2>&1; tail -2 $OUT/pytest_hpair.log
bench() {   # bench <tag> <lib-variant|base> <steps> <warmup>
  lib=$PWD/chemsim_b200/libchemsim_lbm.so; [ $2 != base ] && lib=$PWD/chemsim_b200/libchemsim_lbm_$2.so
  CHEMSIM_LBM_LIB=$lib python bench.py --steps $3 --warmup $4 --no-extras --no-cpu > $OUT/bench_$1.json 2>> $OUT/bench.err
  python -c "import json; d=json.loads(open('$OUT/bench_$1.json').read().strip().splitlines()[-1]); print('$1', round(d['value'],2), d['batch_ms'], d['clocks']['sm_mhz'])"
}
bench drv_base base 20 5
bench drv_hpair hpair 20 5
bench drv_stcs stcs 20 5
bench drv_ty10 s2ty10 20 5
bench sus_base base 200 20
bench sus_hpair hpair 200 20
tail -2 $OUT/bench.err
