#!/bin/bash
# Round 2, GPU call B (N GPUs): multi-GPU parity tests + the driver's bench command at N (+ N=1 first, for the efficiencies)
N=${1:-2}
OUT=gpurun_out/r02b_n$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv
if [ "${SKIP_TESTS:-0}" != "1" ]; then
( time python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $OUT/pytest_multi.log 2>&1
tail -4 $OUT/pytest_multi.log
fi
( time python bench.py --steps 20 --warmup 5 --no-cpu ) > $OUT/bench_n1.json 2> $OUT/bench_n1.err
for halo in p2p nccl; do
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --halo $halo ) > $OUT/bench_n${N}_$halo.json 2> $OUT/bench_n${N}_$halo.err
tail -c 300 $OUT/bench_n${N}_$halo.err
done
for f in $OUT/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'N', d['n_gpus'], round(d['value'],2), 'GLUPS ms/step', round(d['ms_per_step'],4), 'batch', d['batch_ms'], 'reps', d['reps'], d['run']['halo'], d['clocks'])
    for k,v in d.get('extras',{}).items():
        print('   ', k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','efficiency','mass_drift_rel','mass_drift_predicted_rel','wall_s','p2p','nccl','halo','error','aborted','batch_ms')})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
