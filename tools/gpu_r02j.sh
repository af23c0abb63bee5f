#!/bin/bash
# r02j (1 GPU): k_trace_q with per-kind state traffic; occupancy / carve-out sweep; parity re-check
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "wavefront or ray_parity or image_parity or walk" > $O/r02j_tests.log 2>&1; echo "tests rc=$?" >> $O/r02j_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --spp 256 --no-cpu-baseline > $O/r02j_bench_sponza.json 2> $O/r02j_bench_sponza.err
timeout 600 python tools/sweep_env.py LJ_Q_BLOCKS 8,7,6,5,4 --spp 128 > $O/r02j_sweep_blocks.txt 2>&1
timeout 600 python tools/sweep_env.py LJ_Q_CARVEOUT 100,75,65,50 --spp 128 > $O/r02j_sweep_carveout.txt 2>&1
LJ_Q_BLOCKS=6 timeout 600 python tools/sweep_env.py LJ_Q_CARVEOUT 100,65,50,40 --spp 128 > $O/r02j_sweep_carveout_b6.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trace_q' -s 20 -c 2 -o /tmp/r02j_sponza python bench.py --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $O/r02j_ncu.log 2>&1
python tools/ncu_metrics.py /tmp/r02j_sponza.ncu-rep > $O/r02j_sponza_metrics.txt 2>&1
