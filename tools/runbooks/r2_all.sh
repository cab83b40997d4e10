# Round 2, first GPU call (ONE B200, ~10 minutes): everything that was written after round 1's GPU budget
# was spent, in order of importance, each step under its own timeout and in its own process.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/runbooks/r2_all.sh'
set -x
mkdir -p gpurun_out
# 1. the default path: GPU tests, smoke, the bench line (CG now 88 + 64 B/DoF per iteration)
timeout 400 python -m pytest tests -m gpu -q -rfs 2>&1 > gpurun_out/r2_tests.log; tail -15 gpurun_out/r2_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/r2_smoke.log; cat gpurun_out/r2_smoke.log
timeout 300 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cat gpurun_out/r2_bench_default.json
# 2. the gated tests (peer boards on one GPU, TMA line-major tridsol, host batch, swizzled y/z tiles)
for K in peer_boards line_major_tma host_batch yz_rot lineop_tma any_chunk fused_reduction; do
  PBX_TEST_ROUND2=1 timeout 300 python -m pytest tests -m gpu -k $K -q -x 2>&1 | tail -4 > gpurun_out/r2_gated_$K.log
  cat gpurun_out/r2_gated_$K.log
done
# 3. the opt-in variants in the bench line: swizzled y/z tiles, batched end-to-end path
PBX_YZ_ROT=1 PBX_BENCH_E2E_BATCH=1 timeout 300 python bench.py --no-cpu > gpurun_out/r2_bench_rot.json 2> gpurun_out/r2_bench_rot.err
cat gpurun_out/r2_bench_rot.json
PBX_FUSE_TAIL=1 timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/r2_bench_fuse.json 2> gpurun_out/r2_bench_fuse.err
cat gpurun_out/r2_bench_fuse.json
# 4. tridsol batches, both layouts, generic against TMA tiles
timeout 200 python tools/prof_ops.py 256 > gpurun_out/r2_prof_ops.log 2>&1; tail -8 gpurun_out/r2_prof_ops.log
# 5. counters of the y / z passes (planes, bank conflicts, with and without PBX_YZ_ROT)
bash tools/runbooks/r2_zpass.sh
# 6. extents that are not 16 x a power of two: generic kernels against the TMA kernels (PBX_TMA_ANY_T=1)
for anyt in 0 1; do
  PBX_TMA_ANY_T=$anyt timeout 120 python tools/prof_lapl.py --n 384 > gpurun_out/r2_anyt${anyt}_384.log 2>&1; tail -1 gpurun_out/r2_anyt${anyt}_384.log
done
# 7. races: compute-sanitizer racecheck / synccheck over lapl + 20 CG iterations at 64^3 (VERDICT weak 12)
for tool in racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_lapl.py --n 64 --reps 2 --cg-its 20 > gpurun_out/r2_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
