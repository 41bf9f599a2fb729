#!/bin/bash
# Round 2, GPU call E (N GPUs): full -m gpu suite (incl. multi-GPU) + the driver's bench command at N=1 and N
N=${1:-2}
OUT=gpurun_out/r02e_n$N
mkdir -p $OUT
if [ "${SKIP_TESTS:-0}" != "1" ]; then
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log | cut -c1-300
fi
( time python bench.py --steps 20 --warmup 5 --no-cpu ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
for halo in ${HALOS:-p2p nccl}; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --halo $halo ) > $OUT/bench_n${N}_$halo.json 2> $OUT/bench_n${N}_$halo.err
grep -v "^\*\|OMP_NUM\|^$" $OUT/bench_n${N}_$halo.err | tail -5
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'N', d['n_gpus'], round(d['value'],2), 'GLUPS ms/step', round(d['ms_per_step'],4), 'batch', {k: round(v,3) for k,v in d['batch_ms'].items()}, 'reps', d['reps'], d['run']['halo'], d['run']['kernel'], 'launches', d['gpu_launches'], d['clocks'])
    for k,v in d.get('extras',{}).items():
        print('   ', k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','efficiency','mass_drift_residual_rel','wall_s','p2p','nccl','halo','error','aborted')} if isinstance(v, dict) else v)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
