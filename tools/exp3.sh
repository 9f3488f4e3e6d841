#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
TAG="ent-tables rowb=128" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="ent-tables rowb=64" P3DFFT_B200_ROWB=64 python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="ent-tables 512^3" python tools/prof_pair.py --size 512 --pairs 8 --warm 1 2>&1 | tail -1
TAG="ent-tables 256^3" python tools/prof_pair.py --size 256 --pairs 8 --warm 1 2>&1 | tail -1
TAG="single 1024^3" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 --single 2>&1 | tail -1
