timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "tma or fast or full_size" 2>&1 | tail -8
PBX_TMA_YZ=1 timeout 120 python tools/prof_lapl.py --n 512 --reps 4
timeout 120 python tools/prof_lapl.py --n 512 --reps 4
PBX_TMA_YZ=1 timeout 120 python tools/prof_lapl.py --n 256 --reps 4
timeout 120 python tools/prof_lapl.py --n 256 --reps 4
