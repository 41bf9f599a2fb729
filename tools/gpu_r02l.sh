#!/bin/bash
# Round 2, GPU call L (1 GPU): tile-height variants of the two-step kernel
OUT=gpurun_out/r02l
mkdir -p $OUT
for rep in 1 2; do for v in base s2ty8 s2ty32; do
  lib=$PWD/chemsim_b200/libchemsim_lbm.so; [ $v != base ] && lib=$PWD/chemsim_b200/libchemsim_lbm_$v.so
  for dt in f32 f64; do
  CHEMSIM_LBM_LIB=$lib python bench.py --steps 200 --warmup 20 --no-extras --no-cpu --dtype $dt > $OUT/bench_${dt}_${v}_r$rep.json 2>> $OUT/bench.err
  python - $OUT/bench_${dt}_${v}_r$rep.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value'],2), 'GLUPS', d['clocks'])
PY
  done
done; done
CHEMSIM_LBM_LIB=$PWD/chemsim_b200/libchemsim_lbm_s2ty32.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_4096 or ragged or golden" 2>&1 | tail -2
tail -3 $OUT/bench.err
