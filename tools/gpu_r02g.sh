#!/bin/bash
# r02g (1 GPU): full GPU suite, benches of every workload (k_trace_q build), ncu of the Disney shade kernel and the grid-media kernels
# (ncu reports stay on the box: only text summaries come back, gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/handout_parity.jsonl $O/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r02g_tests.log 2>&1; echo "tests rc=$?" >> $O/r02g_tests.log
for W in cbox veach_mi disney_bsdf volpath_test6 vol_cbox_teapot hetvol hetvol_colored; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 > $O/r02g_bench_$W.json 2> $O/r02g_bench_$W.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 12 -c 1 -o /tmp/r02g_disney_shade python bench.py --workload disney_bsdf --steps 1 --warmup 0 --spp 32 --no-cpu-baseline > $O/r02g_ncu_disney.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_flight|k_shade_vol' -s 40 -c 4 -o /tmp/r02g_hetvol python bench.py --workload hetvol_colored --steps 1 --warmup 0 --spp 32 --no-cpu-baseline > $O/r02g_ncu_hetvol.log 2>&1
python tools/ncu_metrics.py /tmp/r02g_disney_shade.ncu-rep /tmp/r02g_hetvol.ncu-rep > $O/r02g_metrics.txt 2>&1
python tools/ncu_lines.py /tmp/r02g_disney_shade.ncu-rep k_shade 70 > $O/r02g_disney_shade_lines.txt 2>&1
python tools/ncu_opcodes.py /tmp/r02g_disney_shade.ncu-rep k_shade 30 > $O/r02g_disney_shade_opcodes.txt 2>&1
python tools/ncu_lines.py /tmp/r02g_hetvol.ncu-rep "k_trace<(int)3>" 50 > $O/r02g_hetvol_trace3_lines.txt 2>&1
python tools/ncu_opcodes.py /tmp/r02g_hetvol.ncu-rep "k_trace<(int)3>" 30 > $O/r02g_hetvol_trace3_opcodes.txt 2>&1
