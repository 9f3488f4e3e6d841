#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -k "fast_kernels or forward_backward or single or many or driver" 2>&1 | tail -3
TAG="il32+uniform" python tools/prof_pair.py --size 1024 --pairs 4 --warm 1 2>&1 | tail -1
TAG="il32+uniform 512^3" python tools/prof_pair.py --size 512 --pairs 8 --warm 1 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_parity.py > gpurun_out/h_mp2.log 2>&1
tail -2 gpurun_out/h_mp2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --grid 2x1 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/h_bench2x1.json 2> gpurun_out/h_bench2.err
