# Round 2, call 7 (`gpurun --gpus 2`): which of (fused exchange, reduction tail) breaks on changing inputs; determinism
# of the segmented kernels on the shape that failed once; the isolated one-GPU peer-board test
set -x
mkdir -p gpurun_out
W=2
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for ZF in 0 1; do for FT in 0 1; do
  PBX_Z_FUSED=$ZF PBX_FUSE_TAIL=$FT run 29555 tools/dist_dyn_check.py 512 128 > gpurun_out/r2g_dyn_zf${ZF}_ft${FT}.log 2>&1
  grep "MatMults\|DYN_CHECK" gpurun_out/r2g_dyn_zf${ZF}_ft${FT}.log | cut -c1-260
done; done
timeout 300 python tools/determinism_check.py > gpurun_out/r2g_determinism.log 2>&1; tail -8 gpurun_out/r2g_determinism.log
timeout 900 python -m pytest tests -m gpu -q -rfs -k "tma_and_generic or fused_reduction or test_cg or peer_boards" > gpurun_out/r2g_tests.log 2>&1; tail -6 gpurun_out/r2g_tests.log
