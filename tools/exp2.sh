#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -k "fast_kernels or forward_backward or single" 2>&1 | tail -5
TAG="new-x rowb=128" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="new-x rowb=64" P3DFFT_B200_ROWB=64 python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="new-x 512^3" python tools/prof_pair.py --size 512 --pairs 8 --warm 1 2>&1 | tail -1
TAG="new-x nopf" P3DFFT_B200_PREFETCH=0 python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
ncu --set full --import-source on --clock-control none -k regex:"xr2c|cstage|xc2r" -c 6 -o gpurun_out/d_prof -f python tools/prof_pair.py --size 1024 --pairs 1 > gpurun_out/d_ncu.log 2>&1
