cd /root/repo
python bench.py --impl reference --steps 20 --warmup 5 --ref-budget 30 | tee gpurun_out/bench_reference.json
python bench.py --steps 20 --warmup 5 | tee gpurun_out/bench_1gpu.json
