cd /root/repo
run() { local n=$1; shift; timeout "${TMO:-600}" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
# scoped synchronisation first, short and bounded: if it fails here everything below runs with world barriers
if TMO=100 run 8 tests/mp_stress.py --iters 300 --grid 2x4 > gpurun_out/mp_stress_8gpu_scoped.log 2>&1; then
  tail -1 gpurun_out/mp_stress_8gpu_scoped.log
else
  echo "SCOPED STRESS FAILED (rc $?) -> world barriers for the rest"; tail -5 gpurun_out/mp_stress_8gpu_scoped.log
  export P3DFFT_B200_SCOPED=0
fi
TMO=200 run 8 tools/ab_multi.py --size 1024 --grid 2x4 --pairs 10 --variants ";P3DFFT_B200_SCOPED=0;P3DFFT_B200_OVERLAP=0;P3DFFT_B200_XSTAGE=0" 2>&1 | grep "^\[\|EXCEPTION" | tee gpurun_out/ab_multi_8gpu_final.log
TMO=240 run 8 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 2 2>&1 | tail -1 | tee gpurun_out/bench_8gpu.json
TMO=200 run 8 bench.py --gpus 8 --nx 2048 --ny 512 --nz 513 --op cheby --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_cheby_8gpu.json
