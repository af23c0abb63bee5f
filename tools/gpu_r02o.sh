#!/bin/bash
# r02o (1 GPU): block-wise common majorant for grid media: tests, benches, steady-state ncu
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "hetvol or volpath or vol_cbox or walk or medium or majorant" > $O/r02o_tests.log 2>&1; echo "tests rc=$?" >> $O/r02o_tests.log
for W in hetvol hetvol_colored vol_cbox_teapot volpath_test6; do
  timeout 400 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > $O/r02o_bench_$W.json 2> $O/r02o_bench_$W.err
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_flight|k_shade_vol' -s 12 -c 4 -o /tmp/r02o_hetvol python bench.py --workload hetvol_colored --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $O/r02o_ncu_hetvol.log 2>&1
python tools/ncu_metrics.py /tmp/r02o_hetvol.ncu-rep > $O/r02o_hetvol_metrics.txt 2>&1
python tools/ncu_lines.py /tmp/r02o_hetvol.ncu-rep "k_trace<(int)3>" 60 > $O/r02o_hetvol_trace3_lines.txt 2>&1
python tools/ncu_lines.py /tmp/r02o_hetvol.ncu-rep "k_flight" 40 > $O/r02o_hetvol_flight_lines.txt 2>&1
python tools/ncu_lines.py /tmp/r02o_hetvol.ncu-rep "k_shade_vol" 40 > $O/r02o_hetvol_shade_lines.txt 2>&1
