#!/bin/bash
# One gpurun call's worth of measurements (1 GPU).  Replaces the per-experiment scratch scripts of round 1.
#   gpurun --timeout 900 -- 'bash tools/gpu_session.sh all'
# Sections: tests | ab | sanitize | ncu | aux        (results under gpurun_out/)
#           mgpu N  -- N-GPU parity and A/B of the multi-GPU switches:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_session.sh mgpu N'
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
what=${1:-all}

if [[ $what == all || $what == tests ]]; then
  python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/tests.log
fi

if [[ $what == all || $what == ab ]]; then
  # A/B of the switchable kernel variants on the headline size: per-stage times from the library's timers
  {
    TAG="default      " python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="R32 two-pass " P3DFFT_B200_R32=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="X tiles of 8 " P3DFFT_B200_XTX8=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="R32 + XTX8   " P3DFFT_B200_R32=1 P3DFFT_B200_XTX8=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="XY pipe G=4  " P3DFFT_B200_XYPIPE=4 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="XY pipe G=4 P" P3DFFT_B200_XYPIPE=4 P3DFFT_B200_XYPIPE_PERSIST=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="XY pipe G=2 P" P3DFFT_B200_XYPIPE=2 P3DFFT_B200_XYPIPE_PERSIST=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="XY pipe G=8  " P3DFFT_B200_XYPIPE=8 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="XY G=4 noring" P3DFFT_B200_XYPIPE=4 P3DFFT_B200_XYPIPE_RING=0 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="bulk stores  " P3DFFT_B200_BULK=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="half-row c2c " P3DFFT_B200_HALF=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="split always " P3DFFT_B200_SPLIT=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="R32 512^3    " P3DFFT_B200_R32=1 python tools/prof_pair.py --size 512 --pairs 12 --warm 2
    TAG="default 512^3" python tools/prof_pair.py --size 512 --pairs 12 --warm 2
    TAG="R32 single   " P3DFFT_B200_R32=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2 --single
    TAG="default singl" python tools/prof_pair.py --size 1024 --pairs 6 --warm 2 --single
  } 2>&1 | tee gpurun_out/ab.log
  # the variants must pass the same parity tests as the defaults
  P3DFFT_B200_XYPIPE=4 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider 2>&1 | tail -3 | tee -a gpurun_out/ab.log
  P3DFFT_B200_BULK=1 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fast_kernels or large" 2>&1 | tail -3 | tee -a gpurun_out/ab.log
  P3DFFT_B200_HALF=1 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fast_kernels or large" 2>&1 | tail -3 | tee -a gpurun_out/ab.log
  P3DFFT_B200_R32=1 P3DFFT_B200_XTX8=1 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fast_kernels or large" 2>&1 | tail -3 | tee -a gpurun_out/ab.log
fi

if [[ $what == all || $what == sanitize ]]; then
  {
    for sz in "64 64 64" "256 128 64" "128 64 1024" "1024 64 128" "64 26 38" "128 22 18"; do   # (the last two: partial X tiles, the case of the r1 prefetch fix)
      echo "== memcheck $sz"; timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
    done
    for sz in "64 64 64" "128 32 1024" "1024 16 64"; do
      echo "== racecheck $sz"; timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
      echo "== racecheck R32 $sz"; P3DFFT_B200_R32=1 timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
    done
  } | tee gpurun_out/sanitizer.log
fi

if [[ $what == all || $what == ncu ]]; then
  ncu --set full --import-source on --clock-control none -k regex:"xr2c|cstage|xc2r" -c 6 -o gpurun_out/stages -f \
      python tools/prof_pair.py --size 1024 --pairs 1 > gpurun_out/ncu_stages.log 2>&1
  P3DFFT_B200_R32=1 ncu --set full --import-source on --clock-control none -k regex:"cstage" -c 4 -o gpurun_out/stages_r32 -f \
      python tools/prof_pair.py --size 1024 --pairs 1 > gpurun_out/ncu_stages_r32.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
fi

if [[ $what == all || $what == aux ]]; then
  python tools/bench_aux.py 1024 | tee gpurun_out/aux_bench.json
fi

if [[ $what == mgpu ]]; then
  N=${2:-2}
  run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
  {
    echo "== parity, default";                 timeout 600 run tests/mp_parity.py 2>&1 | tail -4
    echo "== the reference's own driver binaries and our C drivers on $N GPUs"
    timeout 600 python -m pytest tests/test_zzzz_reference_binaries.py tests/test_c_drivers.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
    echo "== parity, flag barrier";            P3DFFT_B200_FLAGBAR=1 timeout 600 run tests/mp_parity.py 2>&1 | tail -4
    echo "== parity, bulk stores";             P3DFFT_B200_BULK=1 timeout 600 run tests/mp_parity.py 2>&1 | tail -4
    echo "== parity, flag barrier + overlap";  P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4 timeout 600 run tests/mp_parity.py 2>&1 | tail -4
    for env in "" "P3DFFT_B200_FLAGBAR=1" "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4" "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4 P3DFFT_B200_OVERLAP_SMS=40" \
               "P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=8 P3DFFT_B200_OVERLAP_SMS=72" \
               "P3DFFT_B200_BULK=1" "P3DFFT_B200_BULK=1 P3DFFT_B200_FLAGBAR=1"; do
      echo "== bench [$env]"
      env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
          bench.py --gpus "$N" --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1
    done
  } | tee gpurun_out/mgpu_$N.log
fi
