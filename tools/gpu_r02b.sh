#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "wavefront or ray_parity or image_parity or pixel_filter" > $O/r02b_tests.log 2>&1; echo "tests rc=$?" >> $O/r02b_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --spp 256 --no-cpu-baseline > $O/r02b_bench_k1.json 2> $O/r02b_bench_k1.err
timeout 600 python tools/exp_sort.py sponza 21 > $O/r02b_sort.txt 2>&1
timeout 300 python tools/sweep_env.py LJ_Q_REFILL 40,48,56 --spp 128 > $O/r02b_sweep_qrefill.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trace_q' -s 20 -c 2 -o $O/r02b_sponza python bench.py --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $O/r02b_ncu.log 2>&1
python tools/ncu_metrics.py $O/r02b_sponza.ncu-rep > $O/r02b_sponza_metrics.txt 2>&1
