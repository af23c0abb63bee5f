#!/bin/bash
# r02z4 (2 GPUs): NCCL-in-the-process slowdown of k_trace_q<0>, second round of suspects
mkdir -p gpurun_out; O=gpurun_out/r02z4_nccl.txt; : > $O
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline $2 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$1', round(j['value'], 1), 'ms/step', round(j['ms_per_step'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
}
LJ_BENCH_LATE_DIST=1 run "late-init(alloc before NCCL)"
NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1 NCCL_BUFFSIZE=131072 run "1 channel, 128K buffers"
TORCH_NCCL_ENABLE_MONITORING=0 TORCH_NCCL_ASYNC_ERROR_HANDLING=0 TORCH_NCCL_BLOCKING_WAIT=0 run "no watchdog monitoring"
CUDA_MODULE_LOADING=LAZY run "module loading lazy"
CUDA_MODULE_LOADING=EAGER run "module loading eager"
run "no reduce (NCCL initialised, never used)" "--no-reduce"
