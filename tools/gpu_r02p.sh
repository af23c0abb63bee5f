#!/bin/bash
# r02p (1 GPU): scheduling knobs of the walk / flight kernels after the block majorants
mkdir -p gpurun_out; O=gpurun_out/r02p_sweeps.txt; : > $O
python tools/sweep_env.py LJ_TRAV_MIN 4,8,12,16,24 --spp 256 --workload hetvol_colored >> $O 2>&1
python tools/sweep_env.py LJ_TRACK_REFILL 8,16,24,28,32 --spp 256 --workload hetvol_colored >> $O 2>&1
python tools/sweep_env.py LJ_SHADOW_CHUNK 32,64,128,256 --spp 256 --workload hetvol_colored >> $O 2>&1
python tools/sweep_env.py LJ_TRAV_MIN 4,8,16 --spp 256 --workload hetvol >> $O 2>&1
