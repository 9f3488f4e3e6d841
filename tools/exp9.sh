#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
TAG="final" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
p3dfft_b200/lib/wave_roundtrip 256 256 256 1 1 3
