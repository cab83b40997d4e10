# Round 2, second GPU call (`gpurun --gpus 2`): the two gated tests that failed in the first call, the suite with the
# promoted defaults, the new bench legs, and the peer boards over cudaIpc on 2 GPUs.
set -x
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES_ALL=$CUDA_VISIBLE_DEVICES
for K in peer_boards line_major_tma; do
  PBX_TEST_ROUND2=1 timeout 600 python -m pytest tests -m gpu -k $K -q -x --tb=short 2>&1 | tail -60 > gpurun_out/r2b_gated_$K.log
  tail -30 gpurun_out/r2b_gated_$K.log
done
timeout 400 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2b_tests.log 2>&1; tail -8 gpurun_out/r2b_tests.log
timeout 600 python bench.py > gpurun_out/r2b_bench_1.json 2> gpurun_out/r2b_bench_1.err; cat gpurun_out/r2b_bench_1.json; tail -5 gpurun_out/r2b_bench_1.err
N=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$N" -ge 2 ]; then
  W=2
  for PS in 0 1 2; do   # 2: peer boards + reduction tails inside the kernels (PBX_FUSE_TAIL)
    if [ $PS -eq 2 ]; then export PBX_FUSE_TAIL=1; PS=1; TAG=2; else unset PBX_FUSE_TAIL; TAG=$PS; fi
    PBX_PEER_SYNC=$PS timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 \
      --master-port 29555 tools/dist_check.py 256 > gpurun_out/r2b_dist_check_w${W}_ps${TAG}.log 2>&1
    tail -3 gpurun_out/r2b_dist_check_w${W}_ps${TAG}.log
    PBX_PEER_SYNC=$PS PBX_BENCH_MG_SLABS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 \
      --master-port 29556 bench.py --gpus $W --no-cpu --quick > gpurun_out/r2b_bench_w${W}_ps${TAG}.json 2> gpurun_out/r2b_bench_w${W}_ps${TAG}.err
    cat gpurun_out/r2b_bench_w${W}_ps${TAG}.json; tail -3 gpurun_out/r2b_bench_w${W}_ps${TAG}.err
  done
  unset PBX_FUSE_TAIL
fi
