#!/bin/bash
# Round 2, GPU call U (1 GPU): the last build (adjacent-pair phase A + streaming phase-B stores as defaults): parity
# suite, A/B against write-back stores, the driver's bench command, ncu full of the kernel
OUT=gpurun_out/r02u
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
bench() {
  lib=$PWD/chemsim_b200/libchemsim_lbm.so; [ $2 != base ] && lib=$PWD/chemsim_b200/libchemsim_lbm_$2.so
  CHEMSIM_LBM_LIB=$lib python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > $OUT/bench_$1.json 2>> $OUT/bench.err
  python -c "import json; d=json.loads(open('$OUT/bench_$1.json').read().strip().splitlines()[-1]); print('$1', round(d['value'],2), d['batch_ms'], d['clocks']['sm_mhz'])"
}
bench drv_default base
bench drv_nostcs2 nostcs2
python bench.py --steps 20 --warmup 5 > $OUT/bench_driver_cmd.json 2> $OUT/bench_driver_cmd.err
python -c "import json; d=json.loads(open('$OUT/bench_driver_cmd.json').read().strip().splitlines()[-1]); print('driver cmd', round(d['value'],2), d['batch_ms'], d['roofline']['frac'], d['clocks'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], {k: round(v['value'],1) for k,v in d['extras'].items() if 'value' in v})"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:step2_kernel -s 10 -c 1 -f -o $OUT/prof_step2_bgk_f32_last \
  python bench.py --steps 20 --warmup 6 --reps 1 --no-cpu --no-extras > $OUT/ncu.log 2>&1
ls $OUT | wc -l
