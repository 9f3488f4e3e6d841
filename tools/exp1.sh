#!/bin/bash
# experiment: tile row width and gather tile order, 1024^3 double 1x1
cd /root/repo
for cfg in "64 -1" "128 -1" "128 1" "128 8" "128 128" "64 32" ; do
  set -- $cfg
  TAG="rowb=$1 bord=$2" P3DFFT_B200_ROWB=$1 P3DFFT_B200_BORD=$2 python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
done
