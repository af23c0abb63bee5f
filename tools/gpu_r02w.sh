#!/bin/bash
# r02w (1 GPU): pool-size sweep on sponza (does an L2-resident pool pay?) and on hetvol_colored
mkdir -p gpurun_out; O=gpurun_out/r02w_pool.txt; : > $O
for P in 524288 786432 1048576 2097152 4194304 8388608; do
  for W in sponza hetvol_colored; do
  python bench.py --workload $W --steps 2 --warmup 2 --spp 256 --pool $P --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$W pool=$P', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
  done
done
