#!/bin/bash
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 tests/mp_parity.py --grids 2x4 > gpurun_out/m_mp8.log 2>&1
tail -3 gpurun_out/m_mp8.log
$TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/m_bench8.json 2> gpurun_out/m_bench8.err
$TR --master-port 29531 bench.py --gpus 8 --grid 1x8 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/m_bench8_1x8.json 2> gpurun_out/m_bench8_1x8.err
P3DFFT_B200_P2P=0 $TR --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/m_bench8_nccl.json 2> gpurun_out/m_bench8_nccl.err
python tools/p3drun.py -n 8 --port 29650 p3dfft_b200/lib/wave_roundtrip 256 256 256 2 4 2
