# Round 2, call 16 (`gpurun --gpus 2`): 128-plane slabs of 1024^2 lines (segmented y pass on the TMA kernels inside the
# slab path: the per-rank configuration of the 1024^3 weak-scaling point) against one GPU; the bench line on two GPUs
set -x
mkdir -p gpurun_out
W=2
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PBX_CHECK_CG_MAXIT=100 PBX_CHECK_MG=0 run 29555 tools/dist_check.py 1024 256 > gpurun_out/r2p_dist_check_1024.log 2>&1; tail -n 3 gpurun_out/r2p_dist_check_1024.log | cut -c1-300
run 29556 bench.py --gpus $W --no-cpu --quick > gpurun_out/r2p_bench_w$W.json 2> gpurun_out/r2p_bench_w$W.err
grep '^{' gpurun_out/r2p_bench_w$W.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'GDoF/s', d['value'], 'ms', d['ms_per_step'], 'cg s', d['cg']['time_s'], 'its', d['cg']['its'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'])"
