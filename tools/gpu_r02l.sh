#!/bin/bash
# r02l (1 GPU): steady-state ncu captures of the Disney shade pass and the grid-media kernels (earlier captures hit tail waves)
mkdir -p gpurun_out; O=gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 8 -c 2 -o /tmp/r02l_disney python bench.py --workload disney_bsdf --steps 1 --warmup 0 --spp 128 --no-cpu-baseline > $O/r02l_ncu_disney.log 2>&1
python tools/ncu_metrics.py /tmp/r02l_disney.ncu-rep > $O/r02l_disney_metrics.txt 2>&1
python tools/ncu_opcodes.py /tmp/r02l_disney.ncu-rep "k_shade<(int)4, (int)2>" 30 > $O/r02l_disney_opcodes.txt 2>&1
python tools/ncu_lines.py /tmp/r02l_disney.ncu-rep "k_shade<(int)4, (int)2>" 60 > $O/r02l_disney_lines.txt 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_flight|k_shade_vol' -s 12 -c 4 -o /tmp/r02l_hetvol python bench.py --workload hetvol_colored --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $O/r02l_ncu_hetvol.log 2>&1
python tools/ncu_metrics.py /tmp/r02l_hetvol.ncu-rep > $O/r02l_hetvol_metrics.txt 2>&1
python tools/ncu_lines.py /tmp/r02l_hetvol.ncu-rep "k_trace<(int)3>" 60 > $O/r02l_hetvol_trace3_lines.txt 2>&1
python tools/ncu_lines.py /tmp/r02l_hetvol.ncu-rep "k_flight" 40 > $O/r02l_hetvol_flight_lines.txt 2>&1
