#!/bin/bash
# r02a: new parity tests, k_trace_q vs k_trace A/B, refill sweep, steady-state ncu of the queue kernels
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_tests.log 2>&1; echo "tests rc=$?" >> $O/r02a_tests.log
for K in 1 0; do
  LJ_TRACE_KERNEL=$K timeout 300 python bench.py --steps 3 --warmup 3 --spp 256 --no-cpu-baseline > $O/r02a_bench_k$K.json 2> $O/r02a_bench_k$K.err
done
timeout 600 python tools/sweep_env.py LJ_Q_REFILL 24,32,40,48,56,64 --spp 128 > $O/r02a_sweep_qrefill.txt 2>&1
timeout 300 python tools/sweep_env.py LJ_Q_CHUNK 64,128,256,512 --spp 128 > $O/r02a_sweep_qchunk.txt 2>&1
# steady-state wave (skip the scene build + first waves): regen / trace_q<0> / shade / trace_q<1>
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trace_q|k_shade|k_regen' -s 40 -c 4 -o $O/r02a_sponza python bench.py --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $O/r02a_ncu.log 2>&1
python tools/ncu_metrics.py $O/r02a_sponza.ncu-rep > $O/r02a_sponza_metrics.txt 2>&1
nvidia-smi > $O/r02a_smi.txt
