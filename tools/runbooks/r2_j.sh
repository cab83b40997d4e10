# Round 2, call 10 (one B200): L2-order / cache-hint variants on the brick one rank of the 8-GPU run holds
# (512 x 512 x 64), the reordered thin-slab boundary sweep, and the probe variants of the segmented-tile defect
set -x
mkdir -p gpurun_out
L=gpurun_out/r2j_cgbrick.log; : > $L
for v in "PBX_L2_HINTS=0 PBX_UPDR_YFRONT=0" "PBX_L2_HINTS=1 PBX_UPDR_YFRONT=0" "PBX_L2_HINTS=2 PBX_UPDR_YFRONT=0" \
         "PBX_L2_HINTS=3 PBX_UPDR_YFRONT=0" "PBX_L2_HINTS=3 PBX_UPDR_YFRONT=1" "PBX_L2_HINTS=0 PBX_UPDR_YFRONT=1"; do
  env $v timeout 120 python tools/prof_cgbrick.py 512 512 64 >> $L 2>&1
done
env PBX_L2_HINTS=0 timeout 120 python tools/prof_cgbrick.py 512 512 512 60 >> $L 2>&1
env PBX_L2_HINTS=3 timeout 120 python tools/prof_cgbrick.py 512 512 512 60 >> $L 2>&1
env PBX_L2_HINTS=3 PBX_UPDR_YFRONT=1 timeout 120 python tools/prof_cgbrick.py 512 512 512 60 >> $L 2>&1
env PBX_L2_HINTS=0 timeout 120 python tools/prof_cgbrick.py 512 512 128 >> $L 2>&1
env PBX_L2_HINTS=3 PBX_UPDR_YFRONT=1 timeout 120 python tools/prof_cgbrick.py 512 512 128 >> $L 2>&1
cat $L
S=gpurun_out/r2j_slab.log; : > $S
for v in "PBX_SLAB_L2_ORDER=0 PBX_L2_HINTS=0" "PBX_SLAB_L2_ORDER=1 PBX_L2_HINTS=0" "PBX_SLAB_L2_ORDER=1 PBX_L2_HINTS=3"; do
  echo "$v" >> $S; env $v timeout 120 python tools/prof_slab.py 512 8 >> $S 2>&1
done
cat $S
REPS=40 timeout 300 python tools/seg_defect_probe2.py 48,512,1088 > gpurun_out/r2j_seg_probe2.log 2>&1; cat gpurun_out/r2j_seg_probe2.log
timeout 600 python -m pytest tests -m gpu -q -x -k "cg or zslab or lapl" > gpurun_out/r2j_tests.log 2>&1; tail -4 gpurun_out/r2j_tests.log
