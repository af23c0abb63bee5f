#!/bin/bash
# r02m (1 GPU): Disney shade pass over a compacted slot queue
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "disney or homework or golden" > $O/r02m_tests.log 2>&1; echo "tests rc=$?" >> $O/r02m_tests.log
for W in disney_bsdf sponza; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > $O/r02m_bench_$W.json 2> $O/r02m_bench_$W.err
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 8 -c 2 -o /tmp/r02m_disney python bench.py --workload disney_bsdf --steps 1 --warmup 0 --spp 128 --no-cpu-baseline > $O/r02m_ncu_disney.log 2>&1
python tools/ncu_metrics.py /tmp/r02m_disney.ncu-rep > $O/r02m_disney_metrics.txt 2>&1
python tools/ncu_lines.py /tmp/r02m_disney.ncu-rep "k_shade<(int)4, (int)2>" 60 > $O/r02m_disney_lines.txt 2>&1
