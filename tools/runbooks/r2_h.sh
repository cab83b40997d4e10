# Round 2, call 8 (one B200): which pass of (16, 640, 1088) is not deterministic; the suite at HEAD
set -x
mkdir -p gpurun_out
REPS=80 timeout 600 python tools/determinism_check.py 16,640,1088 16,640,64 16,64,1088 32,640,1088 16,512,512 48,640,1088 > gpurun_out/r2h_determinism.log 2>&1; cat gpurun_out/r2h_determinism.log
timeout 900 python -m pytest tests -m gpu -q -rfs > gpurun_out/r2h_tests.log 2>&1; tail -6 gpurun_out/r2h_tests.log
