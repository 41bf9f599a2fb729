#!/bin/bash
# Round 2, GPU call P (2 GPUs): multi-GPU parity suite + the driver's bench command at N=1 and N=2 on the final build
N=${1:-2}
OUT=gpurun_out/r02p_n$N
mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $OUT/pytest_multi.log 2>&1
tail -6 $OUT/pytest_multi.log | cut -c1-300
( time python bench.py --steps 20 --warmup 5 --no-cpu --no-extras ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu ) > $OUT/bench_n${N}_p2p.json 2> $OUT/bench_n${N}_p2p.err
grep -v "^\*\|OMP_NUM\|^$" $OUT/bench_n${N}_p2p.err | tail -5
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
