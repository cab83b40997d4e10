# Round-2 first GPU call for the peer boards (written in round 1 after the GPU budget was spent;
# CPU-harness tested only).  Stage A needs ONE GPU; stage B is the same thing over cudaIpc and needs
# `gpurun --gpus 2` (then 8).  Every stage is wrapped in its own timeout: the kernels' bounded spins
# trap after ~2 minutes if a peer never shows up.
set -x
mkdir -p gpurun_out
# 0. everything that is gated because it has never run on a GPU (peer boards on one GPU, TMA line-major
#    tridsol, host batch), each in its own process so that a trap cannot poison the next
for K in peer_boards line_major_tma host_batch yz_rot lineop_tma any_chunk fused_reduction; do
  PBX_TEST_ROUND2=1 timeout 600 python -m pytest tests -m gpu -k $K -q -x 2>&1 | tail -4 > gpurun_out/r2_gated_$K.log
  cat gpurun_out/r2_gated_$K.log
done
timeout 200 python tools/prof_ops.py 256 > gpurun_out/r2_prof_ops.log 2>&1; tail -8 gpurun_out/r2_prof_ops.log
# A. P slab handles of one process on one GPU, own streams + host threads (tests/test_zslab_gpu.py)
PBX_TEST_PEER_BOARDS=1 timeout 600 python -m pytest tests/test_zslab_gpu.py -k peer_boards -q -x 2>&1 | tail -5 > gpurun_out/r2_peer_onegpu.log
cat gpurun_out/r2_peer_onegpu.log
N=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$N" -ge 2 ]; then
  for W in 2 $( [ "$N" -ge 8 ] && echo 8 ); do
    # B. correctness over NCCL-bootstrapped cudaIpc mappings, with and without the peer boards
    for PS in 0 1 2; do   # 2: peer boards + reduction tails inside the kernels (PBX_FUSE_TAIL)
      if [ $PS -eq 2 ]; then export PBX_FUSE_TAIL=1; PS=1; TAG=2; else unset PBX_FUSE_TAIL; TAG=$PS; fi
      PBX_PEER_SYNC=$PS timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 \
        --master-port 29555 tools/dist_check.py 512 > gpurun_out/r2_dist_check_w${W}_ps${TAG}.log 2>&1
      tail -3 gpurun_out/r2_dist_check_w${W}_ps${TAG}.log
      # C. the bench line (MatMult + CG time-to-solution)
      PBX_PEER_SYNC=$PS PBX_BENCH_MG_SLABS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 \
        --master-port 29556 bench.py --gpus $W --no-cpu --no-e2e > gpurun_out/r2_bench_w${W}_ps${TAG}.json 2> gpurun_out/r2_bench_w${W}_ps${TAG}.err
      cat gpurun_out/r2_bench_w${W}_ps${TAG}.json
    done
    unset PBX_FUSE_TAIL
  done
fi
# D. BASELINE configs[4]: 1024^3 over 8 B200 (1024-point x and y lines, 128-plane slabs), MatMult and a bounded CG
if [ "$N" -ge 8 ]; then
  for PS in 0 1; do
    PBX_PEER_SYNC=$PS timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29557 bench.py --gpus 8 --n 1024 --no-cpu --no-e2e --cg-maxit 300 \
      > gpurun_out/r2_bench_1024_w8_ps${PS}.json 2> gpurun_out/r2_bench_1024_w8_ps${PS}.err
    cat gpurun_out/r2_bench_1024_w8_ps${PS}.json
  done
fi
