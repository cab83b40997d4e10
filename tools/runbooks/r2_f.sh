# Round 2, call 6 (`gpurun --gpus 2`): the fused exchange with staged, tile-major messages against the unfused path
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -rfs -k "tma_and_generic or fused_reduction or test_cg or peer_boards" > gpurun_out/r2f_tests.log 2>&1; tail -4 gpurun_out/r2f_tests.log
W=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for ZF in 1 0; do
  export PBX_Z_FUSED=$ZF PBX_CHECK_CG_MAXIT=300 PBX_CHECK_MG=0
  [ $ZF -eq 1 ] && { run 29555 tools/dist_check.py 512 128 > gpurun_out/r2f_dist_check_zf$ZF.log 2>&1; tail -n 3 gpurun_out/r2f_dist_check_zf$ZF.log; }
  PBX_PROF_PHASES=$((1-ZF)) run 29557 tools/dist_prof.py 512 128 > gpurun_out/r2f_dist_prof_zf$ZF.log 2>&1; tail -n 9 gpurun_out/r2f_dist_prof_zf$ZF.log
done
