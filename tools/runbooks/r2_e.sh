# Round 2, call 5 (`gpurun --gpus 2`): the exchange fused into the z pass on thin slabs (two ranks of 64 planes: what
# each of eight ranks computes at 512^3), against the unfused path; the suite at HEAD; tdma_periodic as three kernels.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2e_tests.log 2>&1; tail -6 gpurun_out/r2e_tests.log
timeout 120 python tools/prof_slab.py 512 8 2>&1 | tail -1 | tee gpurun_out/r2e_prof_slab_8.log
timeout 200 python tools/prof_tdma.py 64 512 2048 > gpurun_out/r2e_prof_tdma.log 2>&1; cat gpurun_out/r2e_prof_tdma.log
W=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for ZF in 0 1; do
  export PBX_Z_FUSED=$ZF PBX_CHECK_CG_MAXIT=300
  run 29555 tools/dist_check.py 512 128 > gpurun_out/r2e_dist_check_zf$ZF.log 2>&1; tail -n 3 gpurun_out/r2e_dist_check_zf$ZF.log
  run 29557 tools/dist_prof.py 512 128 > gpurun_out/r2e_dist_prof_zf$ZF.log 2>&1; tail -n 9 gpurun_out/r2e_dist_prof_zf$ZF.log
done
