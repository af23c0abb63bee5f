#!/bin/bash
# r02i (8 GPUs): strong scaling of the 1024-spp sponza image at N = 8 and 4 (torchrun, as the driver launches it),
# and the same render through the library's own multi-GPU path (lajolla --gpus 8)
mkdir -p gpurun_out; O=gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > $O/r02i_bench_${N}gpu.json 2> $O/r02i_bench_${N}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 5 --warmup 3 --split tiles > $O/r02i_bench_8gpu_tiles.json 2> $O/r02i_bench_8gpu_tiles.err
cd /tmp && for G in 1 8; do
  timeout 300 /root/repo/lajolla_public_b200/lajolla --spp 1024 --gpus $G -o /tmp/sponza_$G.pfm /root/repo/oracle/_ref/scenes/sponza/sponza.xml > /root/repo/gpurun_out/r02i_cli_${G}gpu.log 2>&1
done
timeout 300 python -m pytest tests/test_multigpu.py -q -m gpu > $O/r02i_multigpu_tests.log 2>&1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02i_smi.txt
