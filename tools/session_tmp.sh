cd /root/repo
{ python tools/ab_multi.py --size 1280 --pairs 4 --variants ";P3DFFT_B200_ROWB=64" 2>&1 | grep "^\[\|EXCEPTION"
python tools/ab_multi.py --size 1536 512 1536 --pairs 4 --variants ";P3DFFT_B200_ROWB=64" 2>&1 | grep "^\[\|EXCEPTION"
python tools/ab_multi.py --size 1280 --dtype f32 --pairs 4 --variants ";" 2>&1 | grep "^\[\|EXCEPTION"
} | tee gpurun_out/nonpow2_b.log
python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "non_power_of_two" 2>&1 | tail -2 | tee -a gpurun_out/nonpow2_b.log
