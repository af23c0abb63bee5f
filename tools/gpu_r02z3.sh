#!/bin/bash
# r02z3 (2 GPUs): which NCCL setting keeps k_trace_q<0> at its single-process speed (a process with NCCL initialised ran it 8 % slower even with NVLS off)
mkdir -p gpurun_out; O=gpurun_out/r02z3_nccl.txt; : > $O
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline $2 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$1', round(j['value'], 1), 'ms/step', round(j['ms_per_step'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
}
run "default(NVLS=0)"
NCCL_CUMEM_ENABLE=0 run "CUMEM=0"
NCCL_P2P_DISABLE=1 run "P2P_DISABLE=1"
NCCL_CUMEM_ENABLE=0 NCCL_P2P_DISABLE=1 run "CUMEM=0,P2P_DISABLE=1"
NCCL_SHM_DISABLE=1 NCCL_P2P_DISABLE=1 run "P2P_DISABLE=1,SHM_DISABLE=1(net)"
run "gloo" "--backend gloo"
