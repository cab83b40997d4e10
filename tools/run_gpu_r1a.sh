set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t3.log; cat gpurun_out/t3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cg --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
timeout 120 python tools/prof_lapl.py --n 512 --reps 4 > gpurun_out/prof_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 3 -c 6 -f -o gpurun_out/prof_r1a python tools/prof_lapl.py --n 512 --reps 4 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/prof_plain.log; tail -5 gpurun_out/ncu_full.log
