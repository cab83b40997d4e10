# Round 2, call 12 (`gpurun --gpus 2`, 64-plane slabs of a 512 x 512 x 128 brick = the per-rank work of the 8-GPU run):
# reordered boundary sweep (+ downward walk inside the CG), L2-friendly pass order on slabs, status word instead of the
# per-iteration copy -- against PBX_SLAB_L2_ORDER=0; correctness against one GPU with changing inputs
set -x
mkdir -p gpurun_out
W=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PBX_CHECK_CG_MAXIT=300 PBX_CHECK_MG=0 run 29555 tools/dist_check.py 512 128 > gpurun_out/r2l_dist_check.log 2>&1; tail -n 3 gpurun_out/r2l_dist_check.log
run 29556 tools/dist_dyn_check.py > gpurun_out/r2l_dist_dyn.log 2>&1; tail -n 3 gpurun_out/r2l_dist_dyn.log
run 29557 tools/dist_prof.py 512 128 > gpurun_out/r2l_dist_prof_new.log 2>&1; tail -n 10 gpurun_out/r2l_dist_prof_new.log
PBX_SLAB_L2_ORDER=0 run 29558 tools/dist_prof.py 512 128 > gpurun_out/r2l_dist_prof_order0.log 2>&1; tail -n 10 gpurun_out/r2l_dist_prof_order0.log
run 29559 tools/dist_prof.py 512 > gpurun_out/r2l_dist_prof_512.log 2>&1; tail -n 10 gpurun_out/r2l_dist_prof_512.log
