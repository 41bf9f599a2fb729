#!/bin/bash
# Round 2, GPU call K (2 GPUs): compute-sanitizer memcheck over the sharded parity worker
OUT=gpurun_out/r02k
mkdir -p $OUT
export CHEMSIM_LBM_P2P_TIMEOUT_S=600
for halo in nccl p2p; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 --no-python \
    compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file $OUT/memcheck_${halo}_%p.log \
    python tests/_multigpu_worker.py 1024 70 12 1 f32 $halo > $OUT/worker_$halo.out 2>&1
  echo "halo=$halo rc=$?"; grep -h "MULTIGPU" $OUT/worker_$halo.out; grep -h "ERROR SUMMARY" $OUT/memcheck_${halo}_*.log
done
# single GPU: memcheck + racecheck over a two-step pass with masks and edge tiles
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck_single.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or paint or checkpoint or mirrored or golden" > $OUT/single_memcheck.out 2>&1
echo "single memcheck rc=$?"; tail -2 $OUT/single_memcheck.out; grep -h "ERROR SUMMARY" $OUT/memcheck_single.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck_single.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or test_other_collision" > $OUT/single_racecheck.out 2>&1
echo "single racecheck rc=$?"; tail -2 $OUT/single_racecheck.out; grep -h "RACECHECK SUMMARY" $OUT/racecheck_single.log
