cd /root/repo
for suf in "" _l2128 _l2256 _stcs; do
  echo "== lib suffix '$suf'"
  P3DFFT_B200_LIB_SUFFIX=$suf python tools/ab_multi.py --size 1024 --pairs 8 --variants ";P3DFFT_B200_SPLIT=0" 2>&1 | grep "^\[\|EXCEPTION"
done | tee gpurun_out/ab_1gpu_cachehints.log
