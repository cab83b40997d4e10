# Round 2, call 14 (one B200): rows of p loaded before the z pass's last barrier (fused dot as a template switch);
# z pass against page count and line length; suite
set -x
mkdir -p gpurun_out
L=gpurun_out/r2n_cgbrick.log; : > $L
timeout 200 python tools/prof_cgbrick.py 512 512 64 >> $L 2>&1
timeout 200 python tools/prof_cgbrick.py 512 512 512 60 >> $L 2>&1
timeout 200 python tools/prof_cgbrick.py 512 512 512 60 >> $L 2>&1
timeout 200 python tools/prof_cgbrick.py 256 256 512 60 >> $L 2>&1
timeout 200 python tools/prof_cgbrick.py 512 512 128 60 >> $L 2>&1
timeout 200 python tools/prof_cgbrick.py 1024 512 64 60 >> $L 2>&1
timeout 200 python tools/prof_cgbrick.py 256 512 256 60 >> $L 2>&1
cat $L
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_tests.log 2>&1; tail -n 3 gpurun_out/r2n_tests.log
