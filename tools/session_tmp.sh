cd /root/repo
run() { local n=$1; shift; timeout "${TMO:-600}" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
O="P3DFFT_B200_OVERLAP"
VARS=";P3DFFT_B200_FLAGBAR=0 $O=0;$O=0;P3DFFT_B200_XSTAGE=0;$O=4 ${O}_SMS=56;$O=4 ${O}_SMS=92;$O=2;$O=3;${O}_SHAPE=1,3,3,3,1;${O}_SHAPE=1,2,2,2,1;${O}_SHAPE=2,3,3,2;${O}_SHAPE=1,3,3,3,1 P3DFFT_B200_XSTAGE=0;P3DFFT_B200_R32=0;P3DFFT_B200_BULK=0"
TMO=600 run 4 tools/ab_multi.py --size 1024 --grid 2x2 --pairs 10 --variants "$VARS" 2>&1 | grep "^\[\|EXCEPTION" | tee gpurun_out/ab_multi_4gpu.log
TMO=600 run 4 tests/mp_parity.py 2>&1 | grep -v "^W\|^\[W\|Warning\|warn" | tee gpurun_out/mp_parity_4gpu.log | tail -3
TMO=300 run 4 tests/mp_stress.py --iters 900 2>&1 | tail -1 | tee gpurun_out/mp_stress_4gpu.log
TMO=600 run 4 bench.py --gpus 4 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_4gpu.json
