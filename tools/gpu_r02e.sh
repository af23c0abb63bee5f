#!/bin/bash
# r02e: which part of the 2-rank set-up slows k_trace_q<0>?
mkdir -p gpurun_out; O=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 --no-reduce $EXTRA > $O/r02e_$name.json 2> $O/r02e_$name.err
}
EXTRA="" run control X=1
EXTRA="" run nvls0 NCCL_NVLS_ENABLE=0
EXTRA="" run p2p0 NCCL_P2P_DISABLE=1
EXTRA="" run cumem0 NCCL_CUMEM_ENABLE=0
EXTRA="--backend gloo" run gloo X=1
