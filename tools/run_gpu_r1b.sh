set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t4.log; cat gpurun_out/t4.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cg --no-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -3 gpurun_out/bench3.err; cat gpurun_out/bench3.json
PBX_NO_TMA=1 timeout 120 python tools/prof_lapl.py --n 512 --reps 4
