#!/bin/bash
# r02ze (1 GPU): shared-memory carve-out and refill threshold of k_trace_q, re-swept on the final kernel
mkdir -p gpurun_out; O=gpurun_out/r02ze_sweeps.txt; : > $O
python tools/sweep_env.py LJ_Q_CARVEOUT -,100,80,70,65,60 --spp 256 --workload sponza >> $O 2>&1
python tools/sweep_env.py LJ_Q_REFILL 40,48,56 --spp 256 --workload sponza >> $O 2>&1
