#!/bin/bash
# r02q (1 GPU): k_flight<GRID>, whole primitive group per pass in the walk kernels
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "hetvol or volpath or vol_cbox or walk or medium or majorant" > $O/r02q_tests.log 2>&1; echo "tests rc=$?" >> $O/r02q_tests.log
for W in hetvol hetvol_colored vol_cbox_teapot volpath_test6; do
  timeout 400 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $O/r02q_bench_$W.json 2> $O/r02q_bench_$W.err
done
