cd /root/repo
run() { local n=$1; shift; timeout "${TMO:-600}" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) "$@"; }
V="P3DFFT_B200_FLAGBAR=1"
VARS=";$V"
for c in 4 8; do for s in 56 74 92; do VARS="$VARS;$V P3DFFT_B200_OVERLAP=$c P3DFFT_B200_OVERLAP_SMS=$s"; done; done
VARS="$VARS;$V P3DFFT_B200_OVERLAP=2 P3DFFT_B200_OVERLAP_SMS=74;P3DFFT_B200_OVERLAP=4 P3DFFT_B200_OVERLAP_SMS=74;$V P3DFFT_B200_BULK=0"
TMO=600 run 8 tools/ab_multi.py --size 1024 --grid 2x4 --pairs 10 --variants "$VARS" 2>&1 | grep "^\[\|EXCEPTION" | tee gpurun_out/ab_multi_8gpu.log
TMO=600 run 8 tools/ab_multi.py --size 1024 --grid 1x8 --pairs 10 --variants ";$V;$V P3DFFT_B200_OVERLAP=4 P3DFFT_B200_OVERLAP_SMS=74;$V P3DFFT_B200_OVERLAP=8 P3DFFT_B200_OVERLAP_SMS=74" 2>&1 | grep "^\[\|EXCEPTION" | tee -a gpurun_out/ab_multi_8gpu.log
bash tools/gpu_session.sh mgpu 8 configs
TMO=600 run 8 tests/mp_parity.py --grids 2x4,1x8 2>&1 | grep -v "^W\|^\[W\|Warning\|warn" | tee gpurun_out/mp_parity_8gpu.log | tail -3
P3DFFT_B200_FLAGBAR=1 P3DFFT_B200_OVERLAP=4 TMO=300 run 8 tests/mp_stress.py --iters 600 --grid 2x4 2>&1 | tail -1 | tee gpurun_out/mp_stress_8gpu.log
timeout 300 python -m pytest tests/test_zzzz_reference_binaries.py -m gpu -q -p no:cacheprovider -rA -k "several" 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/refbin_8gpu.log
