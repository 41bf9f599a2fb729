#!/bin/bash
# Round 2, GPU call R (1 GPU): compute-sanitizer over the final build's two-step path (packed additions, L2 prefetch),
# then a last prefetch-distance check in the driver's batch length
OUT=gpurun_out/r02r
mkdir -p $OUT
timeout 140 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_single.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or ragged or paint or other_collision" > $OUT/single_memcheck.out 2>&1
echo "memcheck rc=$?"; tail -2 $OUT/single_memcheck.out; grep -h "ERROR SUMMARY" $OUT/memcheck_single.log
timeout 90 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_single.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > $OUT/single_racecheck.out 2>&1
echo "racecheck rc=$?"; tail -2 $OUT/single_racecheck.out; grep -h "RACECHECK SUMMARY" $OUT/racecheck_single.log
for pf in 96 200; do
  CHEMSIM_LBM_PREFETCH=$pf python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > $OUT/bench_drv_pf$pf.json 2>> $OUT/bench.err
  python -c "import json; d=json.loads(open('$OUT/bench_drv_pf$pf.json').read().strip().splitlines()[-1]); print('pf$pf', round(d['value'],2), d['batch_ms'])"
done
