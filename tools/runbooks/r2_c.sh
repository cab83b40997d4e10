# Round 2, third GPU call (one B200): suite at HEAD, per-rank cost of a 64-plane slab (what one of eight ranks computes),
# the general tridiagonal batches alone and under ncu, and the ncu evidence for the three passes at HEAD.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2c_tests.log 2>&1; tail -6 gpurun_out/r2c_tests.log
for P in 8 4 2; do timeout 120 python tools/prof_slab.py 512 $P 2>&1 | tail -1 | tee gpurun_out/r2c_prof_slab_$P.log; done
timeout 200 python tools/prof_tdma.py 64 512 2048 > gpurun_out/r2c_prof_tdma.log 2>&1; cat gpurun_out/r2c_prof_tdma.log
REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2c_tdma_ncu.csv python tools/prof_tdma.py 64 512 2048 > gpurun_out/r2c_tdma_ncu.log 2>&1
# ncu --set full of one MatMult at HEAD (3 kernels), after the plain run above has exited 0
timeout 120 python tools/prof_lapl.py --n 512 --reps 2 > gpurun_out/r2c_prof_lapl.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 3 -f -o gpurun_out/r2c_full_512 \
  python tools/prof_lapl.py --n 512 --reps 2 > gpurun_out/r2c_full_512.log 2>&1
tail -2 gpurun_out/r2c_prof_lapl.log
# launch list of the default bench command (shares of the kernels in a step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --quick --no-parity --cg-maxit 3 > gpurun_out/r2c_launches.log 2>&1
tail -2 gpurun_out/r2c_launches.log
