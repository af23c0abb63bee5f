#!/bin/bash
# r02u (1 GPU): staged walk (general tracking kernel, auto selection) + two-pass k_shade_vol; A/B against commit 52ded9b on the same box
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "hetvol or volpath or vol_cbox or walk or medium or majorant" > $O/r02u_tests.log 2>&1; echo "tests rc=$?" >> $O/r02u_tests.log
A=$O/r02u_ab.txt; : > $A
for W in hetvol hetvol_colored volpath_test6 vol_cbox_teapot; do
  for L in old new; do
    if [ $L = old ]; then export LJ_LIB=$PWD/lajolla_public_b200/build/libljb200_old.so; else unset LJ_LIB; fi
    python bench.py --workload $W --steps 2 --warmup 2 --spp 256 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); st = j['stage_ms_per_step']
print('$W $L', round(j['value'], 1), {k: round(v, 1) for k, v in st.items()})" >> $A 2>&1
  done
done
unset LJ_LIB
