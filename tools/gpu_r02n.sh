#!/bin/bash
# r02n (1 GPU): local majorants (block-wise tracking) for grid media
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "hetvol or volpath or vol_cbox or walk or medium or majorant" > $O/r02n_tests.log 2>&1; echo "tests rc=$?" >> $O/r02n_tests.log
for W in hetvol hetvol_colored; do
  timeout 400 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $O/r02n_bench_$W.json 2> $O/r02n_bench_$W.err
done
for B in 4 16 0; do
  LJ_MAJ_BLOCK=$B timeout 400 python bench.py --workload hetvol_colored --steps 2 --warmup 3 --spp 256 --no-cpu-baseline > $O/r02n_bench_hetvol_colored_B$B.json 2> $O/r02n_bench_hetvol_colored_B$B.err
  LJ_MAJ_BLOCK=$B timeout 400 python bench.py --workload hetvol --steps 2 --warmup 3 --spp 256 --no-cpu-baseline > $O/r02n_bench_hetvol_B$B.json 2> $O/r02n_bench_hetvol_B$B.err
done
