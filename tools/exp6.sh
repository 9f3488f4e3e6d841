#!/bin/bash
cd /root/repo
ncu --set full --import-source on --clock-control none -k regex:"xr2c|cstage|xc2r" -c 6 -o gpurun_out/i_prof -f python tools/prof_pair.py --size 1024 --pairs 1 > gpurun_out/i_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/i_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
