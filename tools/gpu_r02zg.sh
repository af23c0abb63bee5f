#!/bin/bash
# r02zg (1 GPU): after the entry-face fix of the block majorants: the volpath tests and the two grid-media benches
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "majorant or walk or medium or hetvol or vol_cbox" > $O/r02zg_tests.log 2>&1; echo "tests rc=$?" >> $O/r02zg_tests.log
for W in hetvol hetvol_colored; do
  timeout 300 python bench.py --workload $W --steps 2 --warmup 2 --no-cpu-baseline > $O/r02zg_bench_$W.json 2> $O/r02zg_bench_$W.err
done
