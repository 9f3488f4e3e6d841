#!/bin/bash
cd /root/repo
for sz in "64 64 64" "256 128 64" "128 64 1024" "1024 64 128"; do
  echo "== memcheck $sz"; timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
done
for sz in "64 64 64" "128 32 1024" "1024 16 64"; do
  echo "== racecheck $sz"; timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
done
ncu --set full --import-source on --clock-control none -k regex:"xr2c|cstage|xc2r" -c 6 -o gpurun_out/n_prof -f python tools/prof_pair.py --size 1024 --pairs 1 > gpurun_out/n_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/n_bench_under_ncu.log 2>&1
