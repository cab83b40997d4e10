timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/t5.log; cat gpurun_out/t5.log
timeout 600 python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -3 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1d_ref.json 2> gpurun_out/bench_r1d_ref.err; cat gpurun_out/bench_r1d_ref.json
timeout 120 python tools/prof_lapl.py --n 512 --reps 4 > gpurun_out/prof_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 3 -c 3 -f -o gpurun_out/prof_r1d python tools/prof_lapl.py --n 512 --reps 4 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
timeout 120 python tools/prof_lapl.py --n 512 --reps 2 --cg-its 3 > gpurun_out/prof_plain3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python tools/prof_lapl.py --n 512 --reps 2 --cg-its 3 > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log; nproc; lscpu | grep "Model name"
