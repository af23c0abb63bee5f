#!/bin/bash
# r02t (1 GPU): staged NEE walk (k_walk_begin / k_trace_q<0> on the walk view / k_walk_track / k_walk_finish)
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "hetvol or volpath or vol_cbox or walk or medium or majorant" > $O/r02t_tests.log 2>&1; echo "tests rc=$?" >> $O/r02t_tests.log
A=$O/r02t_ab.txt; : > $A
for W in hetvol hetvol_colored volpath_test6 vol_cbox_teapot; do
  for K in 0 1 2; do
    LJ_WALK_KERNEL=$K python bench.py --workload $W --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$W LJ_WALK_KERNEL=$K', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $A 2>&1
  done
done
