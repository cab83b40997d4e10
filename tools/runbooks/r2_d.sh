# Round 2, 8-GPU call (`gpurun --gpus 8`): correctness of every exchange variant on 8 ranks, the breakdown of the slab
# MatMult, and the bench line (MatMult + CG time-to-solution) per variant.  PS: 0 = NCCL barrier / all-reduces,
# 1 = peer boards, 2 = peer boards + reduction tails inside the kernels.
set -x
mkdir -p gpurun_out
W=${W:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for TAG in 0 2; do
  if [ $TAG -eq 2 ]; then export PBX_FUSE_TAIL=1 PBX_PEER_SYNC=1; else export PBX_FUSE_TAIL=0 PBX_PEER_SYNC=0; fi
  run 29555 tools/dist_check.py 512 > gpurun_out/r2d_dist_check_w${W}_ps${TAG}.log 2>&1; tail -n 2 gpurun_out/r2d_dist_check_w${W}_ps${TAG}.log
  run 29557 tools/dist_prof.py 512 > gpurun_out/r2d_dist_prof_w${W}_ps${TAG}.log 2>&1; tail -n 8 gpurun_out/r2d_dist_prof_w${W}_ps${TAG}.log
  run 29556 bench.py --gpus $W --no-cpu --quick > gpurun_out/r2d_bench_w${W}_ps${TAG}.json 2> gpurun_out/r2d_bench_w${W}_ps${TAG}.err
  grep '^{' gpurun_out/r2d_bench_w${W}_ps${TAG}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'GDoF/s', d['value'], 'ms', d['ms_per_step'], 'cg s', d['cg']['time_s'], 'its', d['cg']['its'], 'launches', d['cg']['gpu_launches'], 'parity', d['parity']['ok'], d['parity']['max_abs_err_over_max_ref'], 'e2e', d['e2e']['value'])"
done
unset PBX_FUSE_TAIL PBX_PEER_SYNC
