#!/bin/bash
# r02d: why is the extend stage 17 % slower in the 2-rank run?  (a) two independent single-GPU processes side by side,
# (b) two ranks without the reduce, (c) two ranks, weak
mkdir -p gpurun_out; O=gpurun_out
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 3 --warmup 3 --spp 512 --no-cpu-baseline > $O/r02d_indep0.json 2> $O/r02d_indep0.err &
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 3 --warmup 3 --spp 512 --no-cpu-baseline > $O/r02d_indep1.json 2> $O/r02d_indep1.err &
wait
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-reduce > $O/r02d_2gpu_noreduce.json 2> $O/r02d_2gpu_noreduce.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 3 --warmup 3 --spp 512 --no-cpu-baseline > $O/r02d_alone0.json 2> $O/r02d_alone0.err
nvidia-smi topo -m > $O/r02d_topo.txt 2>&1
lscpu | head -20 > $O/r02d_cpu.txt
