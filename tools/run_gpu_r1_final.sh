# round-1 closing run on one B200: GPU tests, smoke, the default bench line, then (after each has
# exited 0 without ncu) the ncu launch list of a short bench run and the operator timings
set -x
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/t_r1_final.log; cat gpurun_out/t_r1_final.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/smoke_r1_final.log; cat gpurun_out/smoke_r1_final.log
timeout 300 python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; tail -2 gpurun_out/bench_r1_final2.err; cat gpurun_out/bench_r1_final2.json
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --cg-maxit 3 > gpurun_out/bench_short.json 2>&1 && \
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --cg-maxit 3 > gpurun_out/ncu_list_final.log 2>&1
tail -2 gpurun_out/ncu_list_final.log
timeout 100 python tools/prof_ops.py 256 > gpurun_out/prof_ops_final.log 2>&1; cat gpurun_out/prof_ops_final.log
