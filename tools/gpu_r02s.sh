#!/bin/bash
# Round 2, GPU call S (2 GPUs): compute-sanitizer memcheck over the sharded parity worker on the final build
# (peer-memory halo, two-step slab kernel with packed additions and L2 prefetch)
OUT=gpurun_out/r02s
mkdir -p $OUT
export CHEMSIM_LBM_P2P_TIMEOUT_S=600
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 --no-python \
  compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file $OUT/memcheck_p2p_%p.log \
  python tests/_multigpu_worker.py 1024 70 12 1 f32 p2p > $OUT/worker_p2p.out 2>&1
echo "halo=p2p rc=$?"; grep -h "MULTIGPU" $OUT/worker_p2p.out; grep -h "ERROR SUMMARY" $OUT/memcheck_p2p_*.log
tail -3 $OUT/worker_p2p.out
