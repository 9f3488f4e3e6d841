cd /root/repo
{ python tools/ab_multi.py --size 1024 --pairs 8 --variants ";P3DFFT_B200_R32=0" 2>&1 | grep "^\[\|EXCEPTION"
P3DFFT_B200_UNI=0 python tools/ab_multi.py --size 1024 --pairs 8 --variants ";P3DFFT_B200_R32=0" 2>&1 | grep "^\[\|EXCEPTION"
python tools/ab_multi.py --size 1024 --dtype f32 --pairs 8 --variants ";" 2>&1 | grep "^\[\|EXCEPTION"
P3DFFT_B200_UNI=0 python tools/ab_multi.py --size 1024 --dtype f32 --pairs 8 --variants ";" 2>&1 | grep "^\[\|EXCEPTION"
python tools/ab_multi.py --size 512 --pairs 16 --variants ";" 2>&1 | grep "^\[\|EXCEPTION"
P3DFFT_B200_UNI=0 python tools/ab_multi.py --size 512 --pairs 16 --variants ";" 2>&1 | grep "^\[\|EXCEPTION"
} | tee gpurun_out/ab_1gpu_uni.log
python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "fast_kernels or large or stride1" 2>&1 | tail -2 | tee -a gpurun_out/ab_1gpu_uni.log
