#!/bin/bash
# r02c: strong scaling at N=2 (sponza 1024 spp total), spp split and tile split, + host phase profile
mkdir -p gpurun_out; O=gpurun_out
for SPLIT in spp tiles; do
  LJ_PROFILE_HOST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --split $SPLIT > $O/r02c_bench_2gpu_$SPLIT.json 2> $O/r02c_bench_2gpu_$SPLIT.err
done
# what one rank of an 8-GPU strong run does (128 spp of a 1024-spp image), alone on one GPU: per-render fixed costs
LJ_PROFILE_HOST=1 timeout 300 python bench.py --steps 5 --warmup 3 --spp 128 --no-cpu-baseline > $O/r02c_bench_128spp.json 2> $O/r02c_bench_128spp.err
