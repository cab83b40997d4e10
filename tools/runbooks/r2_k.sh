# Round 2, call 11 (one B200): the proxy fence before every tile-buffer release -- the segmented tiles with and without
# it (many repeats), its cost at 512^3, segmented bricks on the TMA kernels against the generic ones, the whole suite
set -x
mkdir -p gpurun_out
REPS=150 VARIANTS=3,0 timeout 300 python tools/seg_defect_probe2.py 48,512,1088 > gpurun_out/r2k_seg_probe2.log 2>&1
REPS=100 VARIANTS=0 timeout 300 python tools/seg_defect_probe2.py 48,640,1088 >> gpurun_out/r2k_seg_probe2.log 2>&1
cat gpurun_out/r2k_seg_probe2.log
L=gpurun_out/r2k_fence_cost.log; : > $L
for v in 3 0 3 0; do env PBX_YZ_DBG=$v timeout 120 python tools/prof_cgbrick.py 512 512 512 60 >> $L 2>&1; done
for s in "256 512 1024" "128 1024 1024" "64 2048 1024" "64 1024 2048"; do
  for v in 0 1; do env PBX_TMA_SEG=$v timeout 120 python tools/prof_cgbrick.py $s 20 >> $L 2>&1; done
done
cat $L
PBX_TMA_SEG=1 REPS=40 timeout 600 python tools/determinism_check.py 48,640,1088 32,640,1088 16,640,1088 > gpurun_out/r2k_determinism_seg.log 2>&1; cat gpurun_out/r2k_determinism_seg.log
timeout 900 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2k_tests.log 2>&1; tail -4 gpurun_out/r2k_tests.log
PBX_TMA_SEG=1 timeout 900 python -m pytest tests -m gpu -q -x -k "parity or lapl or long" > gpurun_out/r2k_tests_seg.log 2>&1; tail -4 gpurun_out/r2k_tests_seg.log
