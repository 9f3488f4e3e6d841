#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -k "fast_kernels or forward_backward or single or many" 2>&1 | tail -3
TAG="x-rowptr rowb=128" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="x-rowptr 512^3" python tools/prof_pair.py --size 512 --pairs 8 --warm 1 2>&1 | tail -1
TAG="single 1024^3" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 --single 2>&1 | tail -1
