#!/bin/bash
# Runs on an N-GPU box (gpurun --gpus N): BASELINE configs 2 (weak, 8 clips/GPU), 4 (256 clips, strong) and 5 (step sweep) on N ranks.
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
run > gpurun_out/r2_cfg2_n$N.json 2> gpurun_out/r2_cfg2_n$N.err
run --config 4 > gpurun_out/r2_cfg4_n$N.json 2> gpurun_out/r2_cfg4_n$N.err
run --config 5 > gpurun_out/r2_cfg5_n$N.json 2> gpurun_out/r2_cfg5_n$N.err
tail -c 300 gpurun_out/r2_cfg2_n$N.json; tail -c 400 gpurun_out/r2_cfg4_n$N.json | head -c 400; tail -2 gpurun_out/r2_cfg4_n$N.err
