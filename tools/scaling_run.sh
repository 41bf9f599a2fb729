#!/bin/bash
# Multi-GPU scaling measurements (run under gpurun --gpus N). Usage: tools/scaling_run.sh N "<workload:steps> ..."
N=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  wl=${spec%%:*}; steps=${spec##*:}
  out=gpurun_out/scale_${wl}_n${N}.json
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --workload $wl --steps $steps --warmup 10 --no-cpu > $out 2> gpurun_out/scale_${wl}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps $steps --warmup 10 --no-cpu $HALO_ARGS > $out 2> gpurun_out/scale_${wl}_n${N}.err
  fi
  echo "$wl n=$N rc=$?"; tail -1 $out | cut -c 1-420
done
