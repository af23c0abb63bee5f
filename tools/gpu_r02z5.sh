#!/bin/bash
# r02z5 (8 GPUs): strong scaling of the 1024-spp sponza image, N = 1, 2, 4, 8 on one box (allocations before NCCL)
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 2 --no-cpu-baseline > $O/r02z5_bench_1gpu.json 2> $O/r02z5_bench_1gpu.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > $O/r02z5_bench_${N}gpu.json 2> $O/r02z5_bench_${N}gpu.err
done
LJ_BENCH_EARLY_DIST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline > $O/r02z5_bench_8gpu_earlydist.json 2> $O/r02z5_bench_8gpu_earlydist.err
timeout 200 lajolla_public_b200/lajolla --gpus 8 -o /tmp/sponza8.exr oracle/_ref/scenes/sponza/sponza.xml > $O/r02z5_cli_8gpu.log 2>&1
