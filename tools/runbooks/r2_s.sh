# Round 2, call 19 (`gpurun --gpus 4`): the 4-GPU point of the strong-scaling table at HEAD (128-plane slabs)
set -x
mkdir -p gpurun_out
W=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $W --no-cpu --quick > gpurun_out/r2s_bench_w$W.json 2> gpurun_out/r2s_bench_w$W.err
grep '^{' gpurun_out/r2s_bench_w$W.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'GDoF/s', d['value'], 'ms', d['ms_per_step'], 'cg s', d['cg']['time_s'], 'its', d['cg']['its'], 'parity', d['parity']['ok'], 'e2e', d['e2e']['value'], d['config']['rank_sync'])"
