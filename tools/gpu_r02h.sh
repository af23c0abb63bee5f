#!/bin/bash
# r02h (1 GPU): full suite after the class-split shade kernels and the cheaper fp64 terms; regression benches
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/handout_parity.jsonl $O/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r02h_tests.log 2>&1; echo "tests rc=$?" >> $O/r02h_tests.log
for W in disney_bsdf veach_mi sponza cbox; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 > $O/r02h_bench_$W.json 2> $O/r02h_bench_$W.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 60 -c 2 -o /tmp/r02h_disney_shade python bench.py --workload disney_bsdf --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $O/r02h_ncu_disney.log 2>&1
python tools/ncu_metrics.py /tmp/r02h_disney_shade.ncu-rep > $O/r02h_disney_metrics.txt 2>&1
python tools/ncu_opcodes.py /tmp/r02h_disney_shade.ncu-rep "k_shade<(int)4, (int)2>" 25 > $O/r02h_disney_shade_opcodes.txt 2>&1
python tools/ncu_lines.py /tmp/r02h_disney_shade.ncu-rep "k_shade<(int)4, (int)2>" 40 > $O/r02h_disney_shade_lines.txt 2>&1
