#!/bin/bash
# One gpurun call's worth of measurements.  Results under gpurun_out/ (copy what should be judged into profiles/).
#   1 GPU :  gpurun --timeout 900 -- 'bash tools/gpu_session.sh tests|ab|sanitize|ncu|aux|bench|membench|hostpipe'
#   N GPUs:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_session.sh mgpu N [parity] [stress] [ab] [refbin] [bench] [configs]'
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
what=${1:-tests}
run() { local n=$1; shift; timeout "${TMO:-600}" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }

if [[ $what == tests ]]; then
  python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/tests.log
fi

if [[ $what == ab ]]; then
  # per-stage times of the switchable kernel variants on the headline size (library timers)
  {
    TAG="default      " python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="R32 never    " P3DFFT_B200_R32=0 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="R32 always   " P3DFFT_B200_R32=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="bulk stores  " P3DFFT_B200_BULK=1 python tools/prof_pair.py --size 1024 --pairs 6 --warm 2
    TAG="default 512^3" python tools/prof_pair.py --size 512 --pairs 12 --warm 2
    TAG="default singl" python tools/prof_pair.py --size 1024 --pairs 6 --warm 2 --single
  } 2>&1 | tee gpurun_out/ab.log
fi

if [[ $what == sanitize ]]; then
  {
    for sz in "64 64 64" "256 128 64" "128 64 1024" "1024 64 128" "64 26 38" "128 22 18"; do   # (the last two: partial X tiles)
      echo "== memcheck $sz"; timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
    done
    for sz in "64 64 64" "128 32 1024" "1024 16 64"; do
      echo "== racecheck $sz"; timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/prof_pair.py --size $sz --pairs 1 2>&1 | tail -3
    done
  } | tee gpurun_out/sanitizer.log
fi

if [[ $what == ncu ]]; then
  ncu --set full --import-source on --clock-control none -k regex:"xr2c|cstage|xc2r" -c 6 -o gpurun_out/stages -f \
      python tools/prof_pair.py --size 1024 --pairs 1 > gpurun_out/ncu_stages.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-cufft > gpurun_out/bench_under_ncu.log 2>&1
fi

if [[ $what == membench ]]; then
  # access-pattern copies (no FFT arithmetic): the user layout's misaligned rows, flat tiles, store flavours
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu 2>/dev/null
  { ./tools/membench 128; ./tools/membench flat,half; } 2>&1 | tee gpurun_out/membench.log
fi

if [[ $what == hostpipe ]]; then
  python tools/ab_hostpipe.py 2>&1 | tee gpurun_out/ab_hostpipe.log
fi

if [[ $what == aux ]]; then
  python tools/bench_aux.py 1024 | tee gpurun_out/aux_bench.json
fi

if [[ $what == bench ]]; then
  python bench.py --impl reference --steps 20 --warmup 5 | tee gpurun_out/bench_reference.json
  python bench.py --steps 20 --warmup 5 | tee gpurun_out/bench_1gpu.json
fi

if [[ $what == mgpu ]]; then
  N=${2:-2}
  shift 2
  parts=${*:-parity stress ab refbin bench}
  for part in $parts; do
    case $part in
      parity)     # every grid of N ranks: the reference's own matrix against the oracle, rtran_*, r2c_1d, proc queries
        TMO=900 run "$N" tests/mp_parity.py 2>&1 | grep -v "^W\|^\[W\|Warning\|warn" | tee gpurun_out/mp_parity_${N}gpu.log | tail -5 ;;
      stress)     # determinism under a delayed rank, NCCL barrier and flag barrier (+ pipelined tail)
        { TMO=300 run "$N" tests/mp_stress.py --iters 1000 2>&1 | tail -2
          P3DFFT_B200_FLAGBAR=1 TMO=300 run "$N" tests/mp_stress.py --iters 1000 2>&1 | tail -2
          P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4 TMO=300 run "$N" tests/mp_stress.py --iters 1000 2>&1 | tail -2
        } | tee gpurun_out/mp_stress_${N}gpu.log ;;
      ab)         # the multi-GPU switches on the headline size, one job
        TMO=900 run "$N" tools/ab_multi.py --size 1024 2>&1 | grep "^\[\|EXCEPTION" | tee gpurun_out/ab_multi_${N}gpu.log ;;
      refbin)     # the reference's own driver binaries and our C drivers on N GPUs
        timeout 600 python -m pytest tests/test_zzzz_reference_binaries.py tests/test_c_drivers.py -m gpu -q -p no:cacheprovider -rA 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/refbin_${N}gpu.log ;;
      bench)
        TMO=600 run "$N" bench.py --gpus "$N" --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_${N}gpu.json ;;
      configs)    # BASELINE configs 4 and 5 (8 GPUs: 2048^3 single on 1x8 and 2x4; Chebyshev and pruned at the run's grid)
        { if [[ $N == 8 ]]; then
            TMO=600 run 8 bench.py --gpus 8 --nx 2048 --ny 2048 --nz 2048 --dtype f32 --grid 1x8 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1
            TMO=600 run 8 bench.py --gpus 8 --nx 2048 --ny 2048 --nz 2048 --dtype f32 --grid 2x4 --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1
          fi
          TMO=600 run "$N" bench.py --gpus "$N" --nx 2048 --ny 512 --nz 513 --op cheby --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1
          TMO=600 run "$N" bench.py --gpus "$N" --nx 2048 --ny 512 --nz 512 --op pruned --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1
        } | tee gpurun_out/bench_configs_${N}gpu.json ;;
    esac
  done
fi
