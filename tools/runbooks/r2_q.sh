# Round 2, call 17 (`gpurun --gpus 8`): the 1024^3 weak-scaling point of BASELINE configs[4] (2^27 points per GPU, as 512^3
# on one GPU): apply, CG time-to-1e-8, end to end
set -x
mkdir -p gpurun_out
W=8
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29558 bench.py --gpus $W --grid 1024 --no-cpu --quick > gpurun_out/r2q_bench_1024_w$W.json 2> gpurun_out/r2q_bench_1024_w$W.err
grep '^{' gpurun_out/r2q_bench_1024_w$W.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1024^3 N', d['n_gpus'], 'GDoF/s', d['value'], 'ms', d['ms_per_step'], 'cg', d['cg']['time_s'], d['cg']['its'], d['cg']['ms_per_it'], d['cg']['true_residual_rel'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'])"
tail -n 3 gpurun_out/r2q_bench_1024_w$W.err
