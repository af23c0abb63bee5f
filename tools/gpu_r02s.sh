#!/bin/bash
# r02s (1 GPU): A/B on one box -- library of commit 52ded9b (LJ_LIB) against the current one, homogeneous volpath scenes + sponza
mkdir -p gpurun_out; O=gpurun_out/r02s_ab.txt; : > $O
for W in volpath_test6 vol_cbox_teapot sponza veach_mi cbox; do
  for L in old new old new; do
    if [ $L = old ]; then export LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_old.so; else unset LJ_LIB; fi
    python bench.py --workload $W --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$W $L', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $O 2>&1
  done
done
