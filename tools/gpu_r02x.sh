#!/bin/bash
# r02x (1 GPU): pool-size sweep on sponza at the headline spp and at the share of an 8-GPU run
mkdir -p gpurun_out; O=gpurun_out/r02x_pool.txt; : > $O
for S in 1024 128; do
for P in 4194304 8388608 16777216 33554432; do
  python bench.py --workload sponza --steps 2 --warmup 2 --spp $S --pool $P --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('sponza spp=$S pool=$P', round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
done
done
