#!/bin/bash
# Round 2, GPU call H (N GPUs, final build): sharded parity tests, the driver's bench command, config 5 (10k steps,
# mass series), config 4 in f64
N=${1:-8}
OUT=gpurun_out/r02h_n$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv
( time python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "${TESTS_K:-sharded or one_row}" ) > $OUT/pytest_multi.log 2>&1
tail -4 $OUT/pytest_multi.log | cut -c1-300
TORCHRUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time python bench.py --steps 20 --warmup 5 --no-cpu ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
for halo in ${HALOS:-p2p}; do
( time $TORCHRUN bench.py --gpus $N --steps 20 --warmup 5 --halo $halo ) > $OUT/bench_n${N}_$halo.json 2> $OUT/bench_n${N}_$halo.err
done
$TORCHRUN tools/config5_run.py --steps ${C5_STEPS:-10000} > $OUT/config5_n$N.json 2> $OUT/config5.err
$TORCHRUN bench.py --gpus $N --workload strong --dtype f64 --steps 50 --warmup 5 --reps 3 > $OUT/bench_strong_f64_n$N.json 2> $OUT/strong_f64.err
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'N', d['n_gpus'], round(d['value'],2), 'GLUPS ms/step', round(d['ms_per_step'],4), 'batch', {k: round(v,3) for k,v in d['batch_ms'].items()}, 'reps', d['reps'], d['run']['halo'], d['run']['kernel'], d['clocks'])
    for k,v in d.get('extras',{}).items():
        print('   ', k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','efficiency','mass_drift_residual_rel','wall_s','p2p','nccl','halo','error','aborted')} if isinstance(v, dict) else v)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
python - $OUT/config5_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('config5', d['n_gpus'], 'GPUs', round(d['GLUPS'],2), 'GLUPS', d['halo'], d['kernel'])
    for s in d['mass_series']: print('   step', s['step'], 'drift', s['drift_rel'], 'residual', s['residual_rel'])
except Exception as e: print('config5 ERR', e)
PY
tail -3 $OUT/*.err | grep -v "^\*\|OMP_NUM\|^$" | tail -12
