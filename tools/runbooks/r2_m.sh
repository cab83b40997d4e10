# Round 2, call 13 (`gpurun --gpus 2`, 64-plane slabs of a 512 x 512 x 128 brick): raw message planes sent by the y pass,
# slab messages loaded one tile ahead, p prefetched into the L2 for the fused dot -- against PBX_SLAB_YMSG=0; correctness
set -x
mkdir -p gpurun_out
W=2
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PBX_CHECK_CG_MAXIT=300 PBX_CHECK_MG=0 run 29555 tools/dist_check.py 512 128 > gpurun_out/r2m_dist_check.log 2>&1; tail -n 3 gpurun_out/r2m_dist_check.log
run 29556 tools/dist_dyn_check.py 512 128 > gpurun_out/r2m_dist_dyn.log 2>&1; tail -n 3 gpurun_out/r2m_dist_dyn.log
run 29557 tools/dist_prof.py 512 128 > gpurun_out/r2m_dist_prof_new.log 2>&1; grep device gpurun_out/r2m_dist_prof_new.log
PBX_SLAB_YMSG=0 run 29558 tools/dist_prof.py 512 128 > gpurun_out/r2m_dist_prof_ymsg0.log 2>&1; grep device gpurun_out/r2m_dist_prof_ymsg0.log
timeout 200 python tools/prof_cgbrick.py 512 512 64 > gpurun_out/r2m_cgbrick.log 2>&1
timeout 200 python tools/prof_cgbrick.py 512 512 512 60 >> gpurun_out/r2m_cgbrick.log 2>&1; cat gpurun_out/r2m_cgbrick.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_tests.log 2>&1; tail -n 3 gpurun_out/r2m_tests.log
