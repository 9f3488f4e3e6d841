#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -k "fast_kernels or forward_backward or single or many or driver" 2>&1 | tail -3
TAG="cheap-x-tables" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="cheap-x-tables pf=all" P3DFFT_B200_PREFETCH=1000000000 python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="cheap-x-tables 512^3" python tools/prof_pair.py --size 512 --pairs 8 --warm 1 2>&1 | tail -1
TAG="cheap-x-tables 2048x512x512" python tools/prof_pair.py --size 2048 512 512 --pairs 4 --warm 1 2>&1 | tail -1
TAG="single 1024" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 --single 2>&1 | tail -1
