#!/bin/bash
# r02k (1 GPU): render lanes on one GPU -- does overlapping one lane's traversal with the other's shading pay?
mkdir -p gpurun_out; O=gpurun_out/r02k_lanes.txt; : > $O
python tools/exp_lanes.py --lanes 1 --spp 128 >> $O 2>&1
for qb in 8 6 5 4 3; do
  LJ_Q_BLOCKS=$qb python tools/exp_lanes.py --lanes 2 --spp 128 --split 2 >> $O 2>&1
done
LJ_Q_BLOCKS=4 python tools/exp_lanes.py --lanes 2 --spp 128 --split 1 >> $O 2>&1
LJ_Q_BLOCKS=3 python tools/exp_lanes.py --lanes 3 --spp 128 --split 2 >> $O 2>&1
LJ_Q_BLOCKS=2 python tools/exp_lanes.py --lanes 4 --spp 128 --split 2 >> $O 2>&1
LJ_Q_BLOCKS=4 python tools/exp_lanes.py --lanes 2 --spp 128 --split 2 --pool 2097152 >> $O 2>&1
