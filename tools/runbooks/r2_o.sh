# Round 2, call 15 (`gpurun --gpus 8`): the z-slab path at HEAD on eight B200 -- correctness against one GPU, the
# breakdown of the slab apply, the bench line at 512^3 (strong scaling) and the 1024^3 weak-scaling point
set -x
mkdir -p gpurun_out
W=${W:-8}
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
PBX_CHECK_CG_MAXIT=400 run 29555 tools/dist_check.py 512 > gpurun_out/r2o_dist_check_w$W.log 2>&1; tail -n 2 gpurun_out/r2o_dist_check_w$W.log | cut -c1-300
run 29557 tools/dist_prof.py 512 > gpurun_out/r2o_dist_prof_w$W.log 2>&1; grep device gpurun_out/r2o_dist_prof_w$W.log
run 29556 bench.py --gpus $W --no-cpu --quick > gpurun_out/r2o_bench_w$W.json 2> gpurun_out/r2o_bench_w$W.err
grep '^{' gpurun_out/r2o_bench_w$W.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'GDoF/s', d['value'], 'ms', d['ms_per_step'], 'cg s', d['cg']['time_s'], 'its', d['cg']['its'], 'launches', d['cg']['gpu_launches'], 'parity', d['parity']['ok'], d['parity']['max_abs_err_over_max_ref'], 'e2e', d['e2e']['value'])"
run 29558 bench.py --gpus $W --grid 1024 --no-cpu --quick --cg-maxit 300 > gpurun_out/r2o_bench_1024_w$W.json 2> gpurun_out/r2o_bench_1024_w$W.err
grep '^{' gpurun_out/r2o_bench_1024_w$W.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1024^3 N', d['n_gpus'], 'GDoF/s', d['value'], 'ms', d['ms_per_step'], 'cg s', d['cg']['time_s'], 'its', d['cg']['its'], 'ms/it', d['cg']['ms_per_it'], 'parity', d['parity']['ok'])"
tail -n 3 gpurun_out/r2o_bench_1024_w$W.err
