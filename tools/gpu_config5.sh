#!/bin/bash
# usage: bash tools/gpu_config5.sh N   (run under gpurun --gpus N): BASELINE config 5 on N worker peers, 60 s
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/box_config5.py --seconds 60 \
   > gpurun_out/r2k_config5_n$N.json 2> gpurun_out/r2k_config5_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r2k_config5_n$N.err; tail -1 gpurun_out/r2k_config5_n$N.json | cut -c1-1500
